// fp64_probe.cu -- development microbenchmark: per-SM fp64 FMA / division throughput on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_fma(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_div(double* out, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  const double b = 1.0000001;
  for (int i = 0; i < iters; ++i) { a0 = a0 / b; a1 = a1 / b; a2 = a2 / b; a3 = a3 / b; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
__global__ void k_divchain(double* out, int iters) {  // dependent chain: latency of one division
  double a0 = 1.0 + threadIdx.x * 1e-3;
  const double b = 1.0000001;
  for (int i = 0; i < iters; ++i) a0 = a0 / b;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0;
}
__global__ void k_fmachain(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3;
  for (int i = 0; i < iters; ++i) a0 = fma(a0, 1.0000001, 1e-9);
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0;
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 4096;
  for (int threads : {32, 128, 512, 1024}) {
    float ms = timeit([&] { k_fma<<<sms, threads>>>(out, iters); });
    double ops = (double)sms * threads * iters * 8;
    printf("DFMA  threads/SM=%4d: %.3f ms -> %.2f TFLOP/s, %.1f DFMA/clk/SM (at %.2f GHz nominal)\n", threads, ms, 2 * ops / ms / 1e9, ops / (ms * 1e-3) / sms / (ghz * 1e9), ghz);
    ms = timeit([&] { k_div<<<sms, threads>>>(out, iters / 8); });
    ops = (double)sms * threads * (iters / 8) * 4;
    printf("DDIV  threads/SM=%4d: %.3f ms -> %.1f Gdiv/s, %.2f div/clk/SM\n", threads, ms, ops / ms / 1e6, ops / (ms * 1e-3) / sms / (ghz * 1e9));
  }
  float ms = timeit([&] { k_divchain<<<1, 32>>>(out, 4096); });
  printf("DDIV dependent chain: %.1f ns per division (%.0f cycles)\n", ms * 1e6 / 4096, ms * 1e6 / 4096 * ghz);
  ms = timeit([&] { k_fmachain<<<1, 32>>>(out, 65536); });
  printf("DFMA dependent chain: %.2f ns per fma (%.1f cycles)\n", ms * 1e6 / 65536, ms * 1e6 / 65536 * ghz);
  return 0;
}
