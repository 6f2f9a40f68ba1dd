// Micro-benchmark of the FP64 tile GEMM core (pymra_b200/csrc/mra_gemm.cuh) in isolation:
// C(M x N) = A(M x K) B(N x K)^T with one 64x64 tile per CTA.  Prints TF/s per (K, extra smem) case so
// pipeline depth / occupancy choices can be judged against the DMMA peak (tools/fp64_peak.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/gemm_bench tools/gemm_bench.cu
#include <cstdio>
#include <vector>
#include "../pymra_b200/csrc/mra_gemm.cuh"
using namespace mra;

template <int VEC>
__global__ void __launch_bounds__(NT) k_gemm(const double* A, const double* B, double* C, int M, int N, int K) {
  extern __shared__ __align__(16) unsigned char smraw[];
  GemmSmem& gs = *reinterpret_cast<GemmSmem*>(smraw);
  const int m0 = blockIdx.x * TB, n0 = blockIdx.y * TB;
  Acc acc;
  acc.zero();
  auto fa = [&](int rr) -> const double* { return m0 + rr < M ? A + (size_t)(m0 + rr) * K : nullptr; };
  auto fb = [&](int rr) -> const double* { return n0 + rr < N ? B + (size_t)(n0 + rr) * K : nullptr; };
  tile_gemm<VEC, true, true>(acc, K, fa, fb, gs, A);
  tile_epilogue(acc, [&](int row, int col, double v) {
    if (m0 + row < M && n0 + col < N) C[(size_t)(m0 + row) * N + n0 + col] = v;
  });
}

int main() {
  const int M = 148 * 64 * 4, N = 64;
  for (int extra : {0, 34816, 65536}) {
    for (int K : {64, 128, 384, 2048}) {
      double *A, *B, *C;
      cudaMalloc(&A, sizeof(double) * (size_t)M * K);
      cudaMalloc(&B, sizeof(double) * (size_t)N * K);
      cudaMalloc(&C, sizeof(double) * (size_t)M * N);
      std::vector<double> h((size_t)M * K, 0.5);
      cudaMemcpy(A, h.data(), sizeof(double) * (size_t)M * K, cudaMemcpyHostToDevice);
      cudaMemcpy(B, h.data(), sizeof(double) * (size_t)N * K, cudaMemcpyHostToDevice);
      size_t smem = sizeof(GemmSmem) + extra;
      cudaFuncSetAttribute(k_gemm<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      dim3 grid(M / TB, N / TB);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      for (int i = 0; i < 3; ++i) k_gemm<2><<<grid, NT, smem>>>(A, B, C, M, N, K);
      cudaEventRecord(e0);
      const int reps = 10;
      for (int i = 0; i < reps; ++i) k_gemm<2><<<grid, NT, smem>>>(A, B, C, M, N, K);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double c00;
      cudaMemcpy(&c00, C, 8, cudaMemcpyDeviceToHost);
      double tf = 2.0 * M * N * K * reps / (ms * 1e-3) * 1e-12;
      double gbs = 8.0 * ((double)M * K + (double)M * N) * reps / (ms * 1e-3) * 1e-9;
      printf("extra_smem=%6d K=%5d  %.3f ms/launch  %.2f TF/s  %.0f GB/s (A read + C write)  C00=%g (expect %g) err=%s\n",
             extra, K, ms / reps, tf, gbs, c00, 0.25 * K, cudaGetErrorString(cudaGetLastError()));
      cudaFree(A); cudaFree(B); cudaFree(C);
    }
  }
  return 0;
}
