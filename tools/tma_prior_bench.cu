// Stored segment of the prior product, V[tile, K : K+64] = V[tile, 0 : K] * VKL_node^T, for groups of 4 consecutive
// 64-row tiles per CTA (the shape of k_prior_groups at level 6 of BASELINE cfg5), two ways:
//   V0  cp.async (LDGSTS) issued by all 128 threads, cp.async.wait_group + one block barrier per 16-wide chunk
//   V1  TMA: one elected thread issues two cp.async.bulk.tensor.2d per chunk (A box 64 x 16 of V, B box 64 x 16 of VKL,
//       SWIZZLE_128B), completion on a "full" mbarrier per stage; the warps arrive on an "empty" mbarrier per stage, only
//       the issuing thread waits for it -- no block barrier, no per-thread copies or address arithmetic.
//       Fragment loads are bank-conflict free under the TMA swizzle with remapped k slots (lane q of k-step s uses
//       k = 2 s + (q & 1) + 8 (q >> 1), the same for A and B).
// Prints ms and TF/s of both and checks V1 against V0 (bit-identical sums are not expected: the k order differs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_prior_bench tools/tma_prior_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cmath>

constexpr int TB = 64, KC = 16, NT = 128, NSTAGE = 3, PG = 4;

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async_16(double* smem, const double* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ int pos64(int row, int k) { return row * KC + ((((k >> 1) ^ ((row & 3) << 1)) << 1) | (k & 1)); }
__device__ __forceinline__ int pos_tma(int row, int k) { return row * KC + ((((k >> 1) ^ (row & 7)) << 1) | (k & 1)); }

struct Acc {
  double v[2][8][2];
};

// ---------------------------------------------------------------- V0: cp.async
__global__ void __launch_bounds__(NT, 4) k_v0(double* V, long long ldv, const double* VKL, int K, int tiles_per_node) {
  extern __shared__ __align__(1024) double sm[];
  double* sa = sm;
  double* sb = sm + NSTAGE * TB * KC;
  const long long row0 = (long long)blockIdx.x * PG * TB;
  const double* B = VKL + (size_t)(blockIdx.x / (tiles_per_node / PG)) * TB * K;
  const int nk = K / KC, total = nk * PG;
  const int kc = (threadIdx.x & 7) * 2, rb = threadIdx.x >> 3;
  const int lane = threadIdx.x & 31, wm = (threadIdx.x >> 5) * 16, g = lane >> 2, q = lane & 3;
  auto load = [&](int c, int buf) {
    if (c >= total) return;
    const int t = c / nk, k = (c - t * nk) * KC + kc;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = rb + 16 * i;
      cp_async_16(sa + buf * TB * KC + pos64(row, kc), V + (size_t)(row0 + t * TB + row) * ldv + k);
      cp_async_16(sb + buf * TB * KC + pos64(row, kc), B + (size_t)row * K + k);
    }
  };
  for (int s = 0; s < NSTAGE - 1; ++s) {
    load(s, s);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  int buf = 0, c = 0;
  for (int t = 0; t < PG; ++t) {
    Acc acc;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc.v[i][j][0] = acc.v[i][j][1] = 0.0;
    for (int kt = 0; kt < nk; ++kt, ++c) {
      asm volatile("cp.async.wait_group %0;\n" ::"n"(NSTAGE - 2) : "memory");
      __syncthreads();
      int nb = buf + NSTAGE - 1;
      if (nb >= NSTAGE) nb -= NSTAGE;
      load(c + NSTAGE - 1, nb);
      asm volatile("cp.async.commit_group;\n" ::: "memory");
      const double* a = sa + buf * TB * KC;
      const double* b = sb + buf * TB * KC;
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
          for (int j = 0; j < 8; ++j) dmma884(acc.v[i][j], fa[i], fb[j]);
      }
      if (++buf == NSTAGE) buf = 0;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double* vrow = V + (size_t)(row0 + t * TB + wm + i * 8 + g) * ldv + K;
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<double2*>(vrow + j * 8 + q * 2) = make_double2(acc.v[i][j][0], acc.v[i][j][1]);
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// ---------------------------------------------------------------- V1: TMA + mbarrier
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__global__ void __launch_bounds__(NT, 4) k_v1(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                              double* V, long long ldv, int K, int tiles_per_node) {
  extern __shared__ __align__(1024) unsigned char smraw1[];
  double* sa = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smraw1) + 1023) & ~uintptr_t(1023));   // NSTAGE x 8 KB
  double* sb = sa + NSTAGE * TB * KC;
  __shared__ __align__(8) unsigned long long bars[2 * NSTAGE];      // full[0..S), empty[0..S)
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(bars);
  const unsigned sa0 = (unsigned)__cvta_generic_to_shared(sa), sb0 = (unsigned)__cvta_generic_to_shared(sb);
  const long long row0 = (long long)blockIdx.x * PG * TB;
  const int node = blockIdx.x / (tiles_per_node / PG);
  const int nk = K / KC, total = nk * PG;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wm = warp * 16, g = lane >> 2, q = lane & 3;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bar0 + 8 * s, 1);                    // full: one arrive.expect_tx + the bytes
      mbar_init(bar0 + 8 * (NSTAGE + s), 4);         // empty: one arrive per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int c) {                          // one thread: chunk c of the group's stream into stage c % S
    const int s = c % NSTAGE, t = c / nk, k = (c - t * nk) * KC;
    const unsigned full = bar0 + 8 * s;
    mbar_expect_tx(full, 2 * TB * KC * 8);
    tma_load_2d(sa0 + s * TB * KC * 8, &mapA, full, k, (int)(row0 + t * TB));
    tma_load_2d(sb0 + s * TB * KC * 8, &mapB, full, k, node * TB);
  };
  if (threadIdx.x == 0)
    for (int c = 0; c < NSTAGE - 1 && c < total; ++c) issue(c);
  int c = 0;
  for (int t = 0; t < PG; ++t) {
    Acc acc;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc.v[i][j][0] = acc.v[i][j][1] = 0.0;
    for (int kt = 0; kt < nk; ++kt, ++c) {
      const int s = c % NSTAGE;
      // the issuing thread refills the stage of chunk c - 1 (with chunk c + S - 1) once all four warps are done with it
      if (threadIdx.x == 0 && c + NSTAGE - 1 < total) {
        if (c >= 1) mbar_wait(bar0 + 8 * (NSTAGE + (c - 1) % NSTAGE), ((c - 1) / NSTAGE) & 1);
        issue(c + NSTAGE - 1);
      }
      __syncwarp();
      mbar_wait(bar0 + 8 * s, (c / NSTAGE) & 1);
      const double* a = sa + s * TB * KC;
      const double* b = sb + s * TB * KC;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int k = 2 * ks + (q & 1) + 8 * (q >> 1);
        double fa[2], fb[8];
#pragma unroll
        for (int i = 0; i < 2; ++i) fa[i] = a[pos_tma(wm + i * 8 + g, k)];
#pragma unroll
        for (int j = 0; j < 8; ++j) fb[j] = b[pos_tma(j * 8 + g, k)];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) dmma884(acc.v[i][j], fa[i], fb[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8 * (NSTAGE + s));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double* vrow = V + (size_t)(row0 + t * TB + wm + i * 8 + g) * ldv + K;
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<double2*>(vrow + j * 8 + q * 2) = make_double2(acc.v[i][j][0], acc.v[i][j][1]);
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(EncodeFn enc, CUtensorMap* m, void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 8};
  cuuint32_t box[2] = {KC, TB};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
  return r == CUDA_SUCCESS;
}

int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 384;
  const int ngroups = argc > 2 ? atoi(argv[2]) : 16384, tiles_per_node = 16;
  const long long N = (long long)ngroups * PG * TB, ldv = 448;
  const int nnodes = ngroups * PG / tiles_per_node;
  double *V, *VKL;
  if (cudaMalloc(&V, sizeof(double) * N * ldv) != cudaSuccess || cudaMalloc(&VKL, sizeof(double) * (size_t)nnodes * TB * K) != cudaSuccess) {
    printf("cudaMalloc failed\n");
    return 1;
  }
  {
    std::vector<double> h((size_t)1 << 22);
    for (size_t i = 0; i < h.size(); ++i) h[i] = ((i * 2654435761u) % 1000) * 1e-3 - 0.5;
    for (size_t off = 0; off < (size_t)N * ldv; off += h.size())
      cudaMemcpy(V + off, h.data(), sizeof(double) * std::min<size_t>(h.size(), (size_t)N * ldv - off), cudaMemcpyHostToDevice);
    for (size_t off = 0; off < (size_t)nnodes * TB * K; off += h.size())
      cudaMemcpy(VKL + off, h.data(), sizeof(double) * std::min<size_t>(h.size(), (size_t)nnodes * TB * K - off), cudaMemcpyHostToDevice);
  }
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres) != cudaSuccess || !enc) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  CUtensorMap mapA, mapB;
  if (!make_map(enc, &mapA, V, (uint64_t)ldv, (uint64_t)N, (uint64_t)ldv)) return 1;
  if (!make_map(enc, &mapB, VKL, (uint64_t)K, (uint64_t)nnodes * TB, (uint64_t)K)) return 1;
  const size_t smem = sizeof(double) * 2 * NSTAGE * TB * KC + 1024;
  cudaFuncSetAttribute(k_v0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_v1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const double flop = 2.0 * (double)N * TB * K;
  std::vector<double> r0(64 * 8), r1(64 * 8);
  for (int variant = 0; variant < 2; ++variant) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (variant == 0) k_v0<<<ngroups, NT, smem>>>(V, ldv, VKL, K, tiles_per_node);
      else k_v1<<<ngroups, NT, smem>>>(mapA, mapB, V, ldv, K, tiles_per_node);
      cudaEventRecord(e1);
      cudaError_t err = cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("V%d K=%d groups=%d  %.3f ms  %.2f TF/s  %s\n", variant, K, ngroups, ms, flop / (ms * 1e-3) * 1e-12,
             cudaGetErrorString(err != cudaSuccess ? err : cudaGetLastError()));
      if (err != cudaSuccess) return 1;
    }
    // sample: 8 columns of 64 rows spread over the matrix
    std::vector<double>& r = variant == 0 ? r0 : r1;
    for (int i = 0; i < 64; ++i)
      cudaMemcpy(&r[i * 8], V + (size_t)((long long)i * (N / 64) + i) * ldv + K + (i % 8) * 8, 8 * sizeof(double), cudaMemcpyDeviceToHost);
  }
  double maxd = 0, maxv = 0;
  for (size_t i = 0; i < r0.size(); ++i) {
    maxd = std::max(maxd, std::abs(r0[i] - r1[i]));
    maxv = std::max(maxv, std::abs(r0[i]));
  }
  printf("check: max |V0 - V1| = %.3e (max |value| %.3e) -> %s\n", maxd, maxv, maxd <= 1e-11 * std::max(1.0, maxv) ? "OK" : "MISMATCH");
  return 0;
}
