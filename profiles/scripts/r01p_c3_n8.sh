mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --config c3 --steps 10 --warmup 3 > gpurun_out/bench_c3_n8.json 2> gpurun_out/bench_c3_n8.err; tail -c 900 gpurun_out/bench_c3_n8.json; tail -3 gpurun_out/bench_c3_n8.err
