# round 2, 8 GPUs: topology, NCCL exchange tests at world = 8, the north-star bench line (c3 / BH) + c2, CLI on 8 GPUs
O=gpurun_out/r02e; mkdir -p $O
{ nproc; free -g | head -2; nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|model name"; for d in /sys/bus/pci/devices/*; do c=$(cat $d/class 2>/dev/null); if [ "$c" = "0x030200" ]; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done; } > $O/box.txt 2>&1
timeout 900 python -m pytest tests/test_distributed.py "tests/test_host_cli.py::test_cli_on_all_gpus_equals_one_gpu" -m gpu -x -q > $O/pytest_n8.log 2>&1; echo "pytest rc=$?" >> $O/pytest_n8.log; tail -4 $O/pytest_n8.log
for N in 8 2; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N ) > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench N=$N rc=$?"
done
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --config c4 --steps 5 ) > $O/bench_c4_n8.json 2> $O/bench_c4_n8.err; echo "bench c4 N=8 rc=$?"
# the drop-in program over all 8 GPUs: c2, 8e8 rows (70 GB) in /dev/shm
D=/dev/shm/cli_n8; rm -rf $D
./kmdiff_b200/kmdiff-b200 dev-synth $D 10 10 31 4 1000000000 42 1000 800000000 128 matrix
for g in 8 1; do for i in 1 2; do rm -f $D/kmdiff-count.opt; ( time ./kmdiff_b200/kmdiff-b200 diff --km-run $D -1 10 -2 10 -t $(nproc) -c benjamini -o /dev/shm/cli_out_$g -v debug --gpus $g ) 2>&1 | grep -v "partition " > $O/cli_c2_8e8_gpus${g}_run$i.log; done; done
cmp /dev/shm/cli_out_8/case_kmers.fasta /dev/shm/cli_out_1/case_kmers.fasta && echo "8-GPU FASTA == 1-GPU FASTA" >> $O/cli_c2_8e8_gpus8_run2.log
rm -f $D/kmdiff-count.opt; ( time ./oracle/_ref/kmdiff diff --km-run $D -1 10 -2 10 -t $(nproc) -c benjamini -o /dev/shm/ref_out ) > $O/cli_c2_8e8_reference.log 2>&1
rm -rf $D /dev/shm/cli_out_* /dev/shm/ref_out
ls -la $O
