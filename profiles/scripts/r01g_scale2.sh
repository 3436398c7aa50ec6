set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 500 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; tail -c 600 gpurun_out/bench_c2_n2.json; tail -3 gpurun_out/bench_c2_n2.err
timeout 500 $TR bench.py --gpus 2 --config c3 --steps 10 --warmup 3 > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; tail -c 600 gpurun_out/bench_c3_n2.json; tail -3 gpurun_out/bench_c3_n2.err
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; tail -c 400 gpurun_out/bench_ref_n2.json
ls -la gpurun_out
