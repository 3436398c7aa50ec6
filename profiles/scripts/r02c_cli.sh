# where the wall time of `kmdiff-b200 diff` goes: c2, 2e8 rows (17.6 GB) in /dev/shm
mkdir -p gpurun_out/r02c
D=/dev/shm/cli_bd; rm -rf $D
./kmdiff_b200/kmdiff-b200 dev-synth $D 10 10 31 4 1000000000 42 1000 200000000 32 matrix
run() { # name, env...
  name=$1; shift
  for i in 1 2; do rm -f $D/kmdiff-count.opt; ( time env "$@" ./kmdiff_b200/kmdiff-b200 diff --km-run $D -1 10 -2 10 -t 16 -o /dev/shm/cli_out -v debug --gpus 1 ) 2>&1 | grep -v "partition " > gpurun_out/r02c/${name}_run$i.log; rm -rf /dev/shm/cli_out; done
}
run base KMD_TIMING=1
run eager KMD_TIMING=1 CUDA_MODULE_LOADING=EAGER
run t8 KMD_TIMING=1 KMD_READERS=8
run t12 KMD_TIMING=1 KMD_READERS=12
run buf8 KMD_TIMING=1 KMD_BUF_MB=8
run buf32 KMD_TIMING=1 KMD_BUF_MB=32
which strace && {
rm -f $D/kmdiff-count.opt; strace -c -f -o gpurun_out/r02c/strace_summary.txt ./kmdiff_b200/kmdiff-b200 diff --km-run $D -1 10 -2 10 -t 16 -o /dev/shm/cli_out --gpus 1 > /dev/null 2>&1; rm -rf /dev/shm/cli_out
rm -f $D/kmdiff-count.opt; strace -r -o gpurun_out/r02c/strace_main.txt ./kmdiff_b200/kmdiff-b200 diff --km-run $D -1 10 -2 10 -t 16 -o /dev/shm/cli_out --gpus 1 > /dev/null 2>&1; rm -rf /dev/shm/cli_out
awk '$1 > 0.02' gpurun_out/r02c/strace_main.txt > gpurun_out/r02c/strace_main_slow.txt; tail -3 gpurun_out/r02c/strace_main.txt >> gpurun_out/r02c/strace_main_slow.txt; rm gpurun_out/r02c/strace_main.txt
}
python - <<'PY' > gpurun_out/r02c/host_memcpy.txt 2>&1
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
cat gpurun_out/r02c/host_memcpy.txt
rm -rf $D
