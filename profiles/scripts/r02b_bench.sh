# round 2, first measurement: new gpu tests, the bench line (c3 headline + c2), CLI breakdown
mkdir -p gpurun_out/r02b
timeout 900 python -m pytest tests/test_distributed.py tests/test_gpu_parity.py tests/test_host_cli.py tests/test_config1.py -m gpu -x -q > gpurun_out/r02b/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b/pytest_gpu.log
tail -4 gpurun_out/r02b/pytest_gpu.log
( time python bench.py ) > gpurun_out/r02b/bench.json 2> gpurun_out/r02b/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02b/bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02b/bench_ref.json 2> gpurun_out/r02b/bench_ref.err; echo "ref rc=$?"
# CLI breakdown: c2, 2e8 rows, debug log with timestamps
D=/dev/shm/cli_bd; rm -rf $D
./kmdiff_b200/kmdiff-b200 dev-synth $D 10 10 31 4 1000000000 42 1000 200000000 32 matrix
for i in 1 2; do rm -f $D/kmdiff-count.opt; ( time ./kmdiff_b200/kmdiff-b200 diff --km-run $D -1 10 -2 10 -t 16 -o /dev/shm/cli_out_$i -v debug --gpus 1 ) > gpurun_out/r02b/cli_c2_2e8_run$i.log 2>&1; done
rm -f $D/kmdiff-count.opt; ( time ./oracle/_ref/kmdiff diff --km-run $D -1 10 -2 10 -t 16 -o /dev/shm/ref_out ) > gpurun_out/r02b/ref_c2_2e8.log 2>&1
grep -c '>' /dev/shm/cli_out_1/*.fasta /dev/shm/ref_out/*.fasta >> gpurun_out/r02b/ref_c2_2e8.log
rm -rf $D /dev/shm/cli_out_* /dev/shm/ref_out
