// Dense layers of the encoders: fused linear(+concat)(+ReLU)(+group max), row normalisation, and the
// edge formulations of PointConv (set abstraction) and DynamicEdgeConv, all on the exact-fp32 tile GEMM.
#include <algorithm>

#include "gemm.cuh"
#include "kernels.h"

namespace t2p {

// ---------------------------------------------------------------------------------------------------
// linear / concat-linear / group-max linear
// ---------------------------------------------------------------------------------------------------
struct ConcatLoader {
  const float* xa;
  const float* xb;
  int Ka, lda, Kb, ldb, M, row0;
  __device__ __forceinline__ float operator()(int r, int k) const {
    const int row = row0 + r;
    if (row >= M) return 0.f;
    return (k < Ka) ? __ldg(xa + (size_t)row * lda + k) : __ldg(xb + (size_t)row * ldb + (k - Ka));
  }
};

template <bool GROUPMAX>
__global__ void __launch_bounds__(GTHREADS)
linear_kernel(ConcatLoader a, const float* __restrict__ W, const float* __restrict__ bias, int N, int relu,
              float* __restrict__ y, int ldy, int rows_per_group, const int32_t* __restrict__ run_if) {
  __shared__ GemmSmem sm;
  if (run_if != nullptr && *run_if == 0) return;  // conditional re-run (block-uniform)
  a.row0 = blockIdx.x * GBM;
  const int n0 = blockIdx.y * GBN;
  const int K = a.Ka + a.Kb;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  gemm_tile_mainloop(a, W, N, K, n0, acc, sm);

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = n0 + tx * 4 + j;
    bv[j] = (bias != nullptr && col < N) ? __ldg(bias + col) : 0.f;
  }
  if constexpr (!GROUPMAX) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = a.row0 + ty * 4 + i;
      if (row >= a.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + tx * 4 + j;
        if (col >= N) continue;
        float v = acc[i][j] + bv[j];
        if (relu) v = fmaxf(v, 0.f);
        y[(size_t)row * ldy + col] = v;
      }
    }
  } else {
    __shared__ TileReduceSmem rs;
    __shared__ int group[GBM];
    if (threadIdx.x < GBM) {
      const int row = a.row0 + threadIdx.x;
      group[threadIdx.x] = (row < a.M) ? row / rows_per_group : -1;
    }
    float val[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) val[i][j] = fmaxf(acc[i][j] + bv[j], 0.f);
    __syncthreads();
    tile_group_max(val, group, y, ldy, n0, N, rs);
  }
}

static int launch_linear_impl(const float* xa, int Ka, int lda, const float* xb, int Kb, int ldb, int M, const float* W,
                              const float* bias, int N, bool relu, float* y, int ldy, int rows_per_group,
                              cudaStream_t s, const int32_t* run_if = nullptr) {
  if (M <= 0 || N <= 0) return T2P_OK;
  ConcatLoader a{xa, xb ? xb : xa, Ka, lda, Kb, ldb, M, 0};
  dim3 grid((M + GBM - 1) / GBM, (N + GBN - 1) / GBN);
  if (rows_per_group > 0)
    linear_kernel<true><<<grid, GTHREADS, 0, s>>>(a, W, bias, N, 1, y, ldy, rows_per_group, run_if);
  else
    linear_kernel<false><<<grid, GTHREADS, 0, s>>>(a, W, bias, N, relu ? 1 : 0, y, ldy, 0, run_if);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

int launch_linear(const float* x, int M, int K, int ldx, const float* W, const float* bias, int N, bool relu, float* y,
                  int ldy, cudaStream_t s, const int32_t* run_if) {
  return launch_linear_impl(x, K, ldx, nullptr, 0, 0, M, W, bias, N, relu, y, ldy, 0, s, run_if);
}
int launch_linear_concat(const float* xa, int Ka, int lda, const float* xb, int Kb, int ldb, int M, const float* W,
                         const float* bias, int N, bool relu, float* y, int ldy, cudaStream_t s, const int32_t* run_if) {
  return launch_linear_impl(xa, Ka, lda, xb, Kb, ldb, M, W, bias, N, relu, y, ldy, 0, s, run_if);
}
int launch_linear_groupmax(const float* xa, int Ka, int lda, const float* xb, int Kb, int ldb, int M, const float* W,
                           const float* bias, int N, int rows_per_group, float* out, int ldo, cudaStream_t s,
                           const int32_t* run_if) {
  return launch_linear_impl(xa, Ka, lda, xb, Kb, ldb, M, W, bias, N, true, out, ldo, rows_per_group, s, run_if);
}

__global__ void zero_if_kernel(float* __restrict__ p, size_t n, const int32_t* __restrict__ flag) {
  if (*flag == 0) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

int launch_zero_if(float* p, size_t n, const int32_t* flag, cudaStream_t s) {
  if (n == 0) return T2P_OK;
  zero_if_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 1184), 256, 0, s>>>(p, n, flag);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

// ---------------------------------------------------------------------------------------------------
// F.normalize over rows: one warp per row
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2_normalize_rows_kernel(float* __restrict__ x, int M, int width, int ld) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float* p = x + (size_t)row * ld;
  float ss = 0.f;
  for (int c = lane; c < width; c += 32) ss = fmaf(p[c], p[c], ss);
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  for (int c = lane; c < width; c += 32) p[c] *= inv;
}

int launch_l2_normalize_rows(float* x, int M, int width, int ld, cudaStream_t s) {
  if (M <= 0) return T2P_OK;
  l2_normalize_rows_kernel<<<(M + 7) / 8, 256, 0, s>>>(x, M, width, ld);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

// ---------------------------------------------------------------------------------------------------
// edge kernels: A(r,k) = relu(P[rowP[r]][k] + sgn * Q[rowQ[r]][k]), grouped max epilogue
// ---------------------------------------------------------------------------------------------------
struct EdgeLoader {
  const float* P;
  const float* Q;
  const int* rowP;  // smem
  const int* rowQ;  // smem
  const int* grp;   // smem, < 0: row unused
  int ldp, ldq, offp;
  float sgn;
  __device__ __forceinline__ float operator()(int r, int k) const {
    if (grp[r] < 0) return 0.f;
    const float p = __ldg(P + (size_t)rowP[r] * ldp + offp + k);
    const float q = __ldg(Q + (size_t)rowQ[r] * ldq + k);
    return fmaxf(fmaf(sgn, q, p), 0.f);
  }
};

constexpr int SA_MAX_M = 512;

__global__ void __launch_bounds__(GTHREADS)
sa_edge_kernel(const float* __restrict__ T, const float* __restrict__ S, const int32_t* __restrict__ nbr,
               const int32_t* __restrict__ cnt, const int32_t* __restrict__ obj_cell_start, int quirk, int n_obj, int P, int m,
               int C1, const float* __restrict__ W2, const float* __restrict__ b2, int C2, float* __restrict__ out,
               const int32_t* __restrict__ run_if) {
  __shared__ GemmSmem sm;
  __shared__ TileReduceSmem rs;
  __shared__ int incl[SA_MAX_M];
  if (run_if != nullptr && *run_if == 0) return;  // conditional re-run (block-uniform)
  __shared__ int rowT[GBM], rowS[GBM];
  const int t = blockIdx.x, n0 = blockIdx.y * GBN;
  const int tid = threadIdx.x;
  const int extra = quirk ? 1 : 0;

  // objects blockIdx.z, blockIdx.z + gridDim.z, ...: the conditional re-run is launched with a few z-slices only, so that the
  // usual case (flag clear) costs a handful of CTAs instead of one early-exit CTA per (object, tile)
  for (int o = blockIdx.z; o < n_obj; o += gridDim.z) {
  __syncthreads();  // the shared tables of the previous object are no longer read
  for (int c = tid; c < m; c += GTHREADS) incl[c] = cnt[(size_t)o * m + c] + extra;
  __syncthreads();
  for (int off = 1; off < m; off <<= 1) {  // inclusive Hillis-Steele scan, m <= 512
    int v0 = 0, v1 = 0;
    const int i0 = tid, i1 = tid + GTHREADS;
    if (i0 < m && i0 >= off) v0 = incl[i0 - off];
    if (i1 < m && i1 >= off) v1 = incl[i1 - off];
    __syncthreads();
    if (i0 < m) incl[i0] += v0;
    if (i1 < m) incl[i1] += v1;
    __syncthreads();
  }
  const int E = incl[m - 1];
  if (t * GBM >= E) continue;

  if (tid < GBM) {
    const int e = t * GBM + tid;
    int rt = 0, rsv = -1;
    if (e < E) {
      int lo = 0, hi = m - 1;  // smallest c with incl[c] > e
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (incl[mid] > e) hi = mid; else lo = mid + 1;
      }
      const int c = lo;
      const int slot = e - (c ? incl[c - 1] : 0);
      const int cn = cnt[(size_t)o * m + c];
      if (slot < cn) {
        rt = o * P + nbr[((size_t)o * m + c) * T2P_MAX_NEIGHBORS + slot];
      } else {  // flat-index self loop: source = flat point (lo*m + c) of this object's cell
        const int first = obj_cell_start[o];
        const int flat = (o - first) * m + c;
        rt = (first + flat / P) * P + flat % P;
      }
      rsv = o * m + c;
    }
    rowT[tid] = rt;
    rowS[tid] = rsv;
  }
  __syncthreads();

  EdgeLoader a{T, S, rowT, rowS, rowS, C1, C1, 0, -1.f};
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  gemm_tile_mainloop(a, W2, C2, C1, n0, acc, sm);

  const int tx = tid & 15;
  float val[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = n0 + tx * 4 + j;
    const float b = (col < C2) ? __ldg(b2 + col) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) val[i][j] = fmaxf(acc[i][j] + b, 0.f);
  }
  tile_group_max(val, rowS, out, C2, n0, C2, rs);
  }
}

int launch_sa_edge(const float* T, const float* S, const int32_t* nbr, const int32_t* cnt,
                   const int32_t* obj_cell_start, int quirk, int n_obj, int P, int m, int C1, const float* W2,
                   const float* b2, int C2, float* out, cudaStream_t s, const int32_t* run_if) {
  if (n_obj <= 0) return T2P_OK;
  T2P_REQUIRE(m <= SA_MAX_M, T2P_ERR_UNSUPPORTED, "set abstraction: m=%d centres per object > %d", m, SA_MAX_M);
  T2P_REQUIRE(n_obj <= 65535, T2P_ERR_UNSUPPORTED, "set abstraction: n_obj=%d > 65535 per call (chunk the cells)", n_obj);
  const int rows_max = m * (T2P_MAX_NEIGHBORS + (quirk ? 1 : 0));
  dim3 grid((rows_max + GBM - 1) / GBM, (C2 + GBN - 1) / GBN, run_if ? std::min(n_obj, 16) : n_obj);
  sa_edge_kernel<<<grid, GTHREADS, 0, s>>>(T, S, nbr, cnt, obj_cell_start, quirk, n_obj, P, m, C1, W2, b2, C2, out, run_if);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

__global__ void __launch_bounds__(GTHREADS)
edgeconv_kernel(const float* __restrict__ AB, const int32_t* __restrict__ knn, const int32_t* __restrict__ obj_cell,
                int n_obj, int D, const float* __restrict__ W2, const float* __restrict__ b2,
                float* __restrict__ pooled) {
  __shared__ GemmSmem sm;
  __shared__ TileReduceSmem rs;
  __shared__ int rowP[GBM], rowQ[GBM], grp[GBM];
  const int tid = threadIdx.x, n0 = blockIdx.y * GBN;
  if (tid < GBM) {
    const int e = blockIdx.x * GBM + tid;
    const int i = e / T2P_KNN_K;
    int g = -1, j = 0;
    if (i < n_obj) {
      j = knn[e];
      if (j >= 0) g = obj_cell[i];
    }
    rowP[tid] = j < 0 ? 0 : j;
    rowQ[tid] = i < n_obj ? i : 0;
    grp[tid] = g;
  }
  __syncthreads();
  EdgeLoader a{AB, AB, rowP, rowQ, grp, 2 * D, 2 * D, D, 1.f};
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  gemm_tile_mainloop(a, W2, D, D, n0, acc, sm);
  const int tx = tid & 15;
  float val[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = n0 + tx * 4 + j;
    const float b = (col < D) ? __ldg(b2 + col) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) val[i][j] = fmaxf(acc[i][j] + b, 0.f);
  }
  tile_group_max(val, grp, pooled, D, n0, D, rs);
}

int launch_edgeconv(const float* AB, const int32_t* knn, const int32_t* obj_cell, int n_obj, int D, const float* W2,
                    const float* b2, float* pooled, cudaStream_t s) {
  if (n_obj <= 0) return T2P_OK;
  dim3 grid((n_obj * T2P_KNN_K + GBM - 1) / GBM, (D + GBN - 1) / GBN);
  edgeconv_kernel<<<grid, GTHREADS, 0, s>>>(AB, knn, obj_cell, n_obj, D, W2, b2, pooled);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

}  // namespace t2p
