// fp32 CUDA-core tile GEMM used by every dense layer of the encoders (exact-fp32 path).
//
// CTA tile 64 rows x 64 cols, K chunks of 16, 256 threads, 4x4 register tile per thread.  The A operand
// is produced by a functor so that gathers / concatenations / the edge formulation of PointConv and
// EdgeConv ("relu(P[src] +- Q[dst])") are fused into the operand load instead of being materialised in
// HBM; weights are [K,N] row-major (transposed nn.Linear) so that B-tile loads are coalesced float4.
#pragma once
#include "common.cuh"

namespace t2p {

constexpr int GBM = 64;
constexpr int GBN = 64;
constexpr int GBK = 16;
constexpr int GTHREADS = 256;

struct __align__(16) GemmSmem {
  float As[GBK][GBM + 4];  // k-major A tile (pitch 68 floats keeps float4 reads aligned)
  float Bs[GBK][GBN];
};

// acc[i][j] += sum_k A(ty*4+i, k) * W[k][n0 + tx*4 + j]
template <class ALoader>
__device__ __forceinline__ void gemm_tile_mainloop(const ALoader& a, const float* __restrict__ W, int N, int K,
                                                   int n0, float (&acc)[4][4], GemmSmem& sm) {
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int ar = tid >> 2, ak = (tid & 3) * 4;
  const int bk = tid >> 4, bn = (tid & 15) * 4;
  const bool n_vec = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);

  for (int k0 = 0; k0 < K; k0 += GBK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + ak + i;
      sm.As[ak + i][ar] = (k < K) ? a(ar, k) : 0.f;
    }
    {
      const int k = k0 + bk, n = n0 + bn;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K) {
        const float* p = W + (size_t)k * N + n;
        if (n_vec && n + 3 < N) {
          v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          if (n + 0 < N) v.x = __ldg(p + 0);
          if (n + 1 < N) v.y = __ldg(p + 1);
          if (n + 2 < N) v.z = __ldg(p + 2);
          if (n + 3 < N) v.w = __ldg(p + 3);
        }
      }
      *reinterpret_cast<float4*>(&sm.Bs[bk][bn]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&sm.As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&sm.Bs[kk][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
}

// Segmented max over the rows of one 64x64 output tile: values (>= 0, post-ReLU) are staged in smem,
// then one thread per (column, quarter of rows) folds consecutive rows of the same group and issues one
// atomic max per group change.  group[r] < 0 marks rows to skip.
struct __align__(16) TileReduceSmem {
  float v[GBM][GBN + 1];
};

__device__ __forceinline__ void tile_group_max(const float (&val)[4][4], const int* __restrict__ group_smem,
                                               float* __restrict__ out, int ldo, int n0, int N,
                                               TileReduceSmem& rs) {
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) rs.v[ty * 4 + i][tx * 4 + j] = val[i][j];
  __syncthreads();
  const int col = tid & 63, qr = tid >> 6;  // 4 quarters of 16 rows
  if (n0 + col < N) {
    int cur = -1;
    float m = 0.f;
    for (int r = qr * 16; r < qr * 16 + 16; ++r) {
      const int g = group_smem[r];
      if (g != cur) {
        if (cur >= 0) atomic_max_nonneg(out + (size_t)cur * ldo + n0 + col, m);
        cur = g;
        m = 0.f;
      }
      if (g >= 0) m = fmaxf(m, rs.v[r][col]);
    }
    if (cur >= 0) atomic_max_nonneg(out + (size_t)cur * ldo + n0 + col, m);
  }
  __syncthreads();
}

}  // namespace t2p
