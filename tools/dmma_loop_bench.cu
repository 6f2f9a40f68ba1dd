// Inner-loop microbenchmark of the FP64 tile product: operands RESIDENT in shared memory (no global traffic), so
// only the fragment loads (LDS) and the DMMA stream are timed.  Variants:
//   V0  16 x 64 warp tile, one LDS.64 per fragment and k-step (mra_gemm.cuh chunk_mma as shipped in r05a)
//   V1  32 x 32 warp tile (4 + 4 fragment loads per 16 DMMA instead of 2 + 8)
//   V2  16 x 64 warp tile, one LDS.128 per fragment and PAIR of k-steps (k slots remapped: lane q of k-step 2h + e
//       uses k = 8h + 2q + e, the same for A and B, so a 16-byte load feeds two k-steps)
// each with / without a block barrier per 16-wide chunk and at 1..4 CTAs per SM (dynamic smem padding).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/dmma_loop_bench tools/dmma_loop_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int TB = 64, KC = 16, NT = 128, NSTAGE = 3;

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ int pos64(int row, int k) { return row * KC + ((((k >> 1) ^ ((row & 3) << 1)) << 1) | (k & 1)); }
__device__ __forceinline__ int pos128(int row, int c) { return row * KC + ((c ^ ((row & 1) << 2)) << 1); }   // c = 16-byte column

template <int V, bool BAR>
__global__ void __launch_bounds__(NT) k_loop(double* out, int iters) {
  extern __shared__ __align__(16) double sm[];
  double* sa = sm;
  double* sb = sm + NSTAGE * TB * KC;
  for (int i = threadIdx.x; i < NSTAGE * TB * KC; i += NT) {
    sa[i] = 1e-3 * (i % 97);
    sb[i] = 1e-3 * (i % 89);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  double acc[2][8][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  int buf = 0;
  for (int it = 0; it < iters; ++it) {
    if (BAR) __syncthreads();
    const double* a = sa + buf * TB * KC;
    const double* b = sb + buf * TB * KC;
    if (V == 0) {
      const int wm = warp * 16;
#pragma unroll
      for (int ks = 0; ks < KC; ks += 4) {
        double fa[2], fb[8];
#pragma unroll
        for (int i = 0; i < 2; ++i) fa[i] = a[pos64(wm + i * 8 + g, ks + q)];
#pragma unroll
        for (int j = 0; j < 8; ++j) fb[j] = b[pos64(j * 8 + g, ks + q)];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) dmma884(acc[i][j], fa[i], fb[j]);
      }
    } else if (V == 1) {
      const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
#pragma unroll
      for (int ks = 0; ks < KC; ks += 4) {
        double fa[4], fb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) fa[i] = a[pos64(wm + i * 8 + g, ks + q)];
#pragma unroll
        for (int j = 0; j < 4; ++j) fb[j] = b[pos64(wn + j * 8 + g, ks + q)];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i >> 1][(i & 1) * 4 + j], fa[i], fb[j]);
      }
    } else {
      const int wm = warp * 16;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double2 fa[2], fb[8];
#pragma unroll
        for (int i = 0; i < 2; ++i) fa[i] = *reinterpret_cast<const double2*>(a + pos128(wm + i * 8 + g, q + 4 * h));
#pragma unroll
        for (int j = 0; j < 8; ++j) fb[j] = *reinterpret_cast<const double2*>(b + pos128(j * 8 + g, q + 4 * h));
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) dmma884(acc[i][j], fa[i].x, fb[j].x);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) dmma884(acc[i][j], fa[i].y, fb[j].y);
      }
    }
    if (++buf == NSTAGE) buf = 0;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * NT + threadIdx.x] = s;
}

template <int V, bool BAR>
void run(double* out, int ctas_per_sm, int sms) {
  const int iters = 20000;
  const size_t base = sizeof(double) * 2 * NSTAGE * TB * KC;
  size_t smem = base;
  // pad the dynamic shared memory so that exactly ctas_per_sm CTAs fit (227 KB usable per SM, 1 KB reserved per CTA)
  const size_t per = (227 * 1024) / ctas_per_sm - 1024;
  if (per > smem) smem = per & ~size_t(1023);
  cudaFuncSetAttribute(k_loop<V, BAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_loop<V, BAR>, NT, smem);
  const int blocks = sms * occ;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_loop<V, BAR><<<blocks, NT, smem>>>(out, iters);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_loop<V, BAR><<<blocks, NT, smem>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fl = 2.0 * TB * TB * KC * (double)iters * blocks;
  printf("V%d barrier=%d ctas/SM=%d (asked %d)  %.3f ms  %.2f TF/s  err=%s\n", V, (int)BAR, occ, ctas_per_sm, ms,
         fl / (ms * 1e-3) * 1e-12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 8 * NT);
  for (int c : {1, 2, 3, 4}) {
    run<0, false>(out, c, p.multiProcessorCount);
    run<0, true>(out, c, p.multiProcessorCount);
    run<1, false>(out, c, p.multiProcessorCount);
    run<1, true>(out, c, p.multiProcessorCount);
    run<2, false>(out, c, p.multiProcessorCount);
    run<2, true>(out, c, p.multiProcessorCount);
  }
  return 0;
}
