mkdir -p gpurun_out/r02a
{
nproc; free -g | head -2; df -h /tmp /root /dev/shm 2>&1; lscpu | head -25; numactl -H 2>&1 | head -20; nvidia-smi topo -m 2>&1 | head -30;
nvidia-smi --query-gpu=index,pci.bus_id,name --format=csv; for d in /sys/bus/pci/devices/*; do c=$(cat $d/class 2>/dev/null); if [ "$c" = "0x030200" ]; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done
cat /proc/meminfo | head -5; mount | grep -E " / | /tmp " ; python -c "import os; print(os.sched_getaffinity(0))"
} > gpurun_out/r02a/box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a/pytest_gpu.log
tail -5 gpurun_out/r02a/pytest_gpu.log
