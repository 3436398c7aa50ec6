# round 2, 8 GPUs, final tree: NCCL tests at world = 8, multi-GPU CLI tests, the bench line at N = 8 / 4 / 2, c4 at N = 8
O=gpurun_out/r02n; mkdir -p $O
timeout 900 python -m pytest tests/test_distributed.py "tests/test_host_cli.py::test_cli_on_all_gpus_equals_one_gpu" "tests/test_host_cli.py::test_cli_pop_correction_on_all_gpus_equals_one_gpu" -m gpu -q > $O/pytest_n8.log 2>&1; echo "pytest rc=$?" >> $O/pytest_n8.log; tail -4 $O/pytest_n8.log
for N in 8 4 2; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N ) > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench N=$N rc=$?"
done
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus 8 --config c4 --steps 5 ) > $O/bench_c4_n8.json 2> $O/bench_c4_n8.err; echo "bench c4 N=8 rc=$?"
ls -la $O
