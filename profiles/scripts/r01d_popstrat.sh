set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_popstrat.py -q -x -s > gpurun_out/pytest_popstrat.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_popstrat.log; tail -40 gpurun_out/pytest_popstrat.log
