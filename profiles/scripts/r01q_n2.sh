mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_distributed.py -m gpu -x -q > gpurun_out/pytest_nccl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nccl.log; tail -4 gpurun_out/pytest_nccl.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; tail -c 300 gpurun_out/bench_c3_n2.json; tail -2 gpurun_out/bench_c3_n2.err
