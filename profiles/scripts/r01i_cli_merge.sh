set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_cli.py -m gpu -x -q -k "count_files or km_out_dir" > gpurun_out/pytest_cli_merge.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_cli_merge.log; tail -40 gpurun_out/pytest_cli_merge.log
