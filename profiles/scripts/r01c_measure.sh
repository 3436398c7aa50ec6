set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 600 gpurun_out/bench_c2.json
timeout 300 python bench.py --config c3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 300 gpurun_out/bench_c3.json
timeout 300 python bench.py --config c5 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 300 gpurun_out/bench_c5.json
timeout 400 python profiles/sweep_scan.py c3 c5 c4 > gpurun_out/sweep.log 2>&1; tail -50 gpurun_out/sweep.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --rows 200000000 --no-cpu-baseline --e2e-steps 1 --e2e-rows 4194304 > gpurun_out/ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lrt_scan -s 3 -c 2 -o gpurun_out/prof_scan_c3 python bench.py --config c3 --steps 2 --warmup 3 --rows 20000000 --no-cpu-baseline --e2e-steps 1 --e2e-rows 1048576 > gpurun_out/ncu_c3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lrt_scan -s 3 -c 2 -o gpurun_out/prof_scan_c5 python bench.py --config c5 --steps 2 --warmup 3 --rows 80000000 --no-cpu-baseline --e2e-steps 1 --e2e-rows 1048576 > gpurun_out/ncu_c5.log 2>&1
ls -la gpurun_out
