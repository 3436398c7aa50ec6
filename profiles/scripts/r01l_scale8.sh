set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | head -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 400 $TR bench.py --gpus 8 --config c3 --steps 10 --warmup 3 > gpurun_out/bench_c3_n8.json 2> gpurun_out/bench_c3_n8.err; tail -c 700 gpurun_out/bench_c3_n8.json; tail -3 gpurun_out/bench_c3_n8.err
timeout 400 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_c2_n8.json 2> gpurun_out/bench_c2_n8.err; tail -c 500 gpurun_out/bench_c2_n8.json; tail -3 gpurun_out/bench_c2_n8.err
timeout 400 $TR bench.py --gpus 8 --config c5 --steps 10 --warmup 3 > gpurun_out/bench_c5_n8.json 2> gpurun_out/bench_c5_n8.err; tail -c 500 gpurun_out/bench_c5_n8.json; tail -3 gpurun_out/bench_c5_n8.err
