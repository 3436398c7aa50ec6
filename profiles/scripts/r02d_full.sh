# round 2: full GPU suite, CLI timing breakdown, LZ4 decode rates, bench lines, ncu launch list + full captures
O=gpurun_out/r02d; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log

# ---- where the wall time of `kmdiff-b200 diff` goes: c2, 2e8 rows (17.6 GB) in /dev/shm
D=/dev/shm/cli_bd; rm -rf $D
./kmdiff_b200/kmdiff-b200 dev-synth $D 10 10 31 4 1000000000 42 1000 200000000 32 matrix
run() { name=$1; shift
  for i in 1 2; do rm -f $D/kmdiff-count.opt; ( time env "$@" ./kmdiff_b200/kmdiff-b200 diff --km-run $D -1 10 -2 10 -t 16 -o /dev/shm/cli_out -v debug --gpus 1 ) 2>&1 | grep -v "partition " > $O/cli_${name}_run$i.log; rm -rf /dev/shm/cli_out; done; }
run base KMD_TIMING=1
run eager KMD_TIMING=1 CUDA_MODULE_LOADING=EAGER
run t8 KMD_TIMING=1 KMD_READERS=8
run t12 KMD_TIMING=1 KMD_READERS=12
run buf4 KMD_TIMING=1 KMD_BUF_MB=4
run buf32 KMD_TIMING=1 KMD_BUF_MB=32
rm -f $D/kmdiff-count.opt; ( time ./oracle/_ref/kmdiff diff --km-run $D -1 10 -2 10 -t 16 -o /dev/shm/ref_out ) > $O/cli_reference.log 2>&1; rm -rf /dev/shm/ref_out
which strace > /dev/null && { rm -f $D/kmdiff-count.opt; strace -f -c -o $O/strace_summary.txt ./kmdiff_b200/kmdiff-b200 diff --km-run $D -1 10 -2 10 -t 16 -o /dev/shm/cli_out --gpus 1 > /dev/null 2>&1; rm -rf /dev/shm/cli_out; }
rm -rf $D
python - <<'PY' > $O/host_memcpy.txt 2>&1
import numpy as np, time, threading
n=1<<30
src=[np.ones(n//8,dtype=np.uint64) for _ in range(16)]
dst=[np.zeros(n//8,dtype=np.uint64) for _ in range(16)]
for t in (1,4,8,16):
    def w(i): np.copyto(dst[i],src[i])
    th=[threading.Thread(target=w,args=(i,)) for i in range(t)]
    t0=time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]; dt=time.perf_counter()-t0
    print(t,"threads",t*n/dt/1e9,"GB/s copy")
PY

# ---- f1: host LZ4 decode rates
timeout 600 python profiles/scripts/r02_lz4_decode.py 32000000 > $O/lz4_decode.json 2> $O/lz4_decode.err

# ---- bench lines
( time python bench.py ) > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
timeout 600 python bench.py --config c2m > $O/bench_c2m.json 2> $O/bench_c2m.err; echo "c2m rc=$?"
timeout 600 python bench.py --config c4 --steps 5 > $O/bench_c4.json 2> $O/bench_c4.err; echo "c4 rc=$?"
timeout 600 python bench.py --config c5 --steps 5 --no-cli > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 rc=$?"

# ---- ncu: launch list of the bench step, full captures of the scan at c3 / c2 and of the merge pass
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv python bench.py --config c3 --steps 2 --warmup 3 --rows 20000000 --e2e-steps 1 --e2e-rows 262144 --no-cpu-baseline > $O/ncu_list_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lrt_scan_kernel -s 3 -c 1 -o $O/prof_scan_c3 python bench.py --config c3 --steps 1 --warmup 3 --rows 20000000 --e2e-steps 1 --e2e-rows 262144 --no-cpu-baseline > $O/ncu_scan_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lrt_scan_kernel -s 3 -c 1 -o $O/prof_scan_c2 python bench.py --config c2 --steps 1 --warmup 3 --rows 200000000 --e2e-steps 1 --e2e-rows 262144 --no-cpu-baseline > $O/ncu_scan_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:merge_pass_kernel -s 20 -c 2 -o $O/prof_merge_c2m python bench.py --config c2m --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_merge.log 2>&1
ls -la $O
