// scratch/dfma_bench.cu — measurement tool, not product code.
// DFMA pipe of one B200 SM: dependent-issue latency and throughput as a function of the number of
// resident warps and of the independent chains per thread (ILP).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/dfma_bench scratch/dfma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double *out, double a, double b, int iters, long long *cycles) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm, double *out, long long *dcy) {
  const int iters = 2000;
  int threads = 128, blocks_per_sm = warps_per_sm / 4;
  if (warps_per_sm < 4) { threads = 32 * warps_per_sm; blocks_per_sm = 1; }
  const int nb = 148 * blocks_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma<ILP><<<nb, threads>>>(out, 0.999999, 1e-7, 10, dcy);
  cudaEventRecord(e0);
  k_dfma<ILP><<<nb, threads>>>(out, 0.999999, 1e-7, iters, dcy);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cy; cudaMemcpy(&cy, dcy, 8, cudaMemcpyDeviceToHost);
  const double nf = (double)iters * 16 * ILP;              // DFMA per thread
  const double per_sm_clk = nf * threads * blocks_per_sm / (double)cy;   // DFMA lanes / clk / SM
  printf("ILP %2d warps/SM %2d : %8.3f ms  %10lld cycles  %6.2f DFMA/clk/SM  cycles per dependent step %6.2f  TFLOP/s %6.2f\n", ILP, warps_per_sm, ms, cy,
         per_sm_clk, (double)cy / (iters * 16.0), 2.0 * nf * threads * nb / (ms * 1e-3) * 1e-12);
}

int main() {
  double *out; long long *dcy;
  cudaMalloc(&out, sizeof(double) * 148 * 16 * 128);
  cudaMalloc(&dcy, 8);
  const int ws[] = {1, 4, 8, 12, 16, 24, 32};
  for (int w : ws) run<1>(w, out, dcy);
  for (int w : ws) run<2>(w, out, dcy);
  for (int w : ws) run<4>(w, out, dcy);
  for (int w : ws) run<8>(w, out, dcy);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
