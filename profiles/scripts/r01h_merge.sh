set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_merge.py -m gpu -x -q > gpurun_out/pytest_gpu_merge.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_merge.log; tail -30 gpurun_out/pytest_gpu_merge.log
timeout 600 python bench.py --config c4 --steps 5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 900 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
