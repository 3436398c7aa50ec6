set -x
mkdir -p gpurun_out
timeout 600 python bench.py --config c2m --steps 5 > gpurun_out/bench_c2m.json 2> gpurun_out/bench_c2m.err; tail -c 2500 gpurun_out/bench_c2m.json; tail -5 gpurun_out/bench_c2m.err
timeout 300 python bench.py --config c2m --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_c2m.json 2> gpurun_out/bench_ref_c2m.err; tail -c 600 gpurun_out/bench_ref_c2m.json; tail -3 gpurun_out/bench_ref_c2m.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2m.csv python bench.py --config c2m --steps 1 --warmup 3 --rows 2000000 --no-cpu-baseline > gpurun_out/ncu_c2m_list.log 2>&1; tail -2 gpurun_out/ncu_c2m_list.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:merge_pass --launch-skip 8 -c 2 -o gpurun_out/prof_merge_pass python bench.py --config c2m --steps 1 --warmup 3 --rows 2000000 --no-cpu-baseline > gpurun_out/ncu_merge.log 2>&1; tail -2 gpurun_out/ncu_merge.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:glm_irls --launch-skip 5 -c 1 -o gpurun_out/prof_glm_irls python bench.py --config c4 --steps 1 --warmup 3 --rows 6000000 --e2e-steps 1 --e2e-rows 262144 > gpurun_out/ncu_glm.log 2>&1; tail -2 gpurun_out/ncu_glm.log
ls -la gpurun_out
