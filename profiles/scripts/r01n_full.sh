set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 500 $TR bench.py --gpus 2 --config c3 --steps 10 --warmup 3 > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; tail -c 400 gpurun_out/bench_c3_n2.json; tail -3 gpurun_out/bench_c3_n2.err
timeout 600 python profiles/scripts/cli_e2e.py 32000000 > gpurun_out/cli_e2e.jsonl 2> gpurun_out/cli_e2e.err; cat gpurun_out/cli_e2e.jsonl; tail -3 gpurun_out/cli_e2e.err
