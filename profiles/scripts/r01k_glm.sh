set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_popstrat.py -m gpu -x -q > gpurun_out/pytest_gpu_popstrat.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_popstrat.log; tail -15 gpurun_out/pytest_gpu_popstrat.log
timeout 600 python bench.py --config c4 --steps 5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 700 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout 600 python bench.py --config c2m --steps 5 > gpurun_out/bench_c2m.json 2> gpurun_out/bench_c2m.err; head -c 300 gpurun_out/bench_c2m.json; tail -3 gpurun_out/bench_c2m.err
timeout 900 python -m pytest tests/test_host_cli.py -m gpu -x -q -k "count_files or km_out_dir or resume" > gpurun_out/pytest_cli2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_cli2.log; tail -15 gpurun_out/pytest_cli2.log
