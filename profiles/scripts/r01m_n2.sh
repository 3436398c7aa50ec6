set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_distributed.py -m gpu -x -q > gpurun_out/pytest_nccl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nccl.log; tail -15 gpurun_out/pytest_nccl.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 500 $TR bench.py --gpus 2 --config c3 --steps 10 --warmup 3 > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; tail -c 600 gpurun_out/bench_c3_n2.json; tail -3 gpurun_out/bench_c3_n2.err
timeout 500 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; head -c 400 gpurun_out/bench_c2_n2.json; tail -3 gpurun_out/bench_c2_n2.err
timeout 600 python bench.py --config c4 --steps 5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 700 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
