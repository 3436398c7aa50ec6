mkdir -p gpurun_out
timeout 70 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_c2_final.err; head -c 250 gpurun_out/bench_c2_final.json; tail -2 gpurun_out/bench_c2_final.err
