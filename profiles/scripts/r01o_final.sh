set -x
mkdir -p gpurun_out
timeout 200 python profiles/scripts/cli_breakdown.py 8000000 1 > gpurun_out/cli_breakdown.json 2> gpurun_out/cli_breakdown.err; python -c "
import json; j=json.load(open('gpurun_out/cli_breakdown.json')); print({k:v for k,v in j.items() if k!='log'}); print('\n'.join(j['log']))"; tail -3 gpurun_out/cli_breakdown.err
timeout 300 python -m pytest tests/test_gpu_popstrat.py tests/test_host_cli.py -m gpu -x -q -k "count_files or kff or resume or pop_correction" > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log; tail -8 gpurun_out/pytest_final.log
timeout 300 python bench.py --config c4 --steps 5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 500 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
