set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 600 python -m pytest tests/test_gpu_popstrat.py -m gpu -x -q > gpurun_out/pytest_gpu_popstrat.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_popstrat.log; tail -5 gpurun_out/pytest_gpu_popstrat.log
timeout 600 python bench.py --config c4 --steps 5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 1800 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:grm_umma -c 2 -o gpurun_out/prof_grm_c4 python bench.py --config c4 --steps 1 --warmup 3 --rows 6000000 --e2e-steps 1 --e2e-rows 262144 > gpurun_out/ncu_grm.log 2>&1; tail -3 gpurun_out/ncu_grm.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4.csv python bench.py --config c4 --steps 1 --warmup 3 --rows 6000000 --e2e-steps 1 --e2e-rows 262144 > gpurun_out/ncu_c4_list.log 2>&1
ls -la gpurun_out
