mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 60 python -m pytest tests/test_host_cli.py -m gpu -x -q -k "matches_oracle and (bonferroni or benjamini)" > gpurun_out/pytest_cli3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_cli3.log; tail -3 gpurun_out/pytest_cli3.log
