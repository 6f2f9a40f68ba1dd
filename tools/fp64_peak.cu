// FP64 throughput microbenchmark for B200: DMMA (mma.sync m8n8k4) vs DFMA issue rate.
// Prints achieved TFLOP/s.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters) {
  double c[NACC][2];
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters) {
  double c[NACC];
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-7;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double* out; cudaMalloc(&out, sizeof(double) * 148 * 16 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int cps : {1, 2}) {
      int blocks = p.multiProcessorCount * cps, threads = warps * 32 / cps;
      if (threads < 32) continue;
      float ms = timeit([&] { k_dmma<8><<<blocks, threads>>>(out, iters); });
      double fl = 2.0 * 256 * 8 * (double)iters * blocks * (threads / 32);
      float ms16 = timeit([&] { k_dmma<16><<<blocks, threads>>>(out, iters); });
      double fl16 = 2.0 * 256 * 16 * (double)iters * blocks * (threads / 32);
      float msf = timeit([&] { k_dfma<16><<<blocks, threads>>>(out, iters); });
      double flf = 2.0 * 16 * (double)iters * blocks * threads;
      printf("warps/SM=%2d ctas/SM=%d  DMMA(8 acc) %.2f TF/s  DMMA(16 acc) %.2f TF/s  DFMA %.2f TF/s\n", warps, cps,
             fl / ms * 1e-9, fl16 / ms16 * 1e-9, flf / msf * 1e-9);
    }
  }
  return 0;
}
