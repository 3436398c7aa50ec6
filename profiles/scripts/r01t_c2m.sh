mkdir -p gpurun_out
timeout 50 python bench.py --config c2m --steps 3 --warmup 3 --rows 4000000 --no-cpu-baseline > gpurun_out/bench_c2m_final.json 2> gpurun_out/bench_c2m_final.err; head -c 200 gpurun_out/bench_c2m_final.json; tail -2 gpurun_out/bench_c2m_final.err
