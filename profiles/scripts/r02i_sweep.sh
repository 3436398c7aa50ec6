O=gpurun_out/r02i; mkdir -p $O
timeout 600 python -m pytest "tests/test_host_cli.py::test_cli_reports_device_errors_and_input_errors_without_hanging" -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -5 $O/pytest.log
SWEEP_FULL=1 timeout 900 python profiles/sweep_scan.py c3 > $O/sweep_c3.log 2>&1; tail -3 $O/sweep_c3.log
cp gpurun_out/sweep_scan.json $O/sweep_scan_c3.json 2>/dev/null
# CUDA start-up anatomy: cuInit / primary context / first kernel, 5 fresh processes
cat > /tmp/init.cu <<'CU'
#include <chrono>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
__global__ void k() {}
int main() {
  using clk = std::chrono::steady_clock;
  auto t0 = clk::now();
  cuInit(0);
  auto t1 = clk::now();
  cudaSetDevice(0); cudaFree(0);
  auto t2 = clk::now();
  k<<<1, 1>>>(); cudaDeviceSynchronize();
  auto t3 = clk::now();
  void* p; cudaMallocHost(&p, 64 << 20);
  auto t4 = clk::now();
  auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  printf("cuInit %.1f ms, primary context %.1f ms, first kernel %.1f ms, cudaMallocHost(64 MiB) %.1f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4));
  return 0;
}
CU
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/init /tmp/init.cu -lcuda 2>/dev/null
for i in 1 2 3 4 5; do ( time /tmp/init ) 2>&1 | grep -E "cuInit|real" | tr '\n' ' '; echo; sleep 1; done > $O/cuda_init.txt
cat $O/cuda_init.txt
