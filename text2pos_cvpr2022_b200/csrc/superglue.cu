// (a8-a10) SuperGlue head as ONE persistent kernel: one CTA per sample keeps both descriptor sets in
// shared memory through all attentional-GNN layers, the final projection, the log-space Sinkhorn
// iterations and the mutual-nearest-neighbour matching.  Weights ([K,N] transposed, BN folded) are
// streamed from L2; nothing but the inputs and the final outputs touches HBM.
//
// Both sides of a layer share the weights (models/superglue.py:144), so every projection runs once over
// the R = M + N stacked rows; rows [0,M) are the objects (desc0), rows [M,R) the hints (desc1).
#include "kernels.h"

namespace t2p {

constexpr int SG_THREADS = 256;
constexpr int SG_RT = 24;  // rows per register chunk of the small GEMM
constexpr int SG_HEADS = 4;

// Y[r][c] (+)= sum_k Xa[r][k] W[k][c] + sum_k Xb[r][k] W[Ka+k][c] + bias[c], rows [0,RP) with RP % SG_RT == 0.
// One thread per (output column, slice of RT rows of a 24-row chunk), its rows in registers; W coalesced from global/L2 with
// 16 loads in flight per thread.  The per-element arithmetic (k ascending, one fma chain) does not depend on the split.
template <int RT>
__device__ __forceinline__ void sg_gemm_rows(const float* __restrict__ Xa, int Ka, const float* __restrict__ Xb, int Kb, int ldx,
                                             int RP, const float* __restrict__ W, const float* __restrict__ bias, int Nout,
                                             float* __restrict__ Y, int ldy, bool relu, int c, int row_off) {
  const float b = bias ? __ldg(bias + c) : 0.f;
  for (int r0 = row_off; r0 < RP; r0 += SG_RT) {
    float acc[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r] = b;
    for (int seg = 0; seg < 2; ++seg) {
      const float* X = seg ? Xb : Xa;
      const int K = seg ? Kb : Ka;
      const float* Wp = W + (size_t)(seg ? Ka : 0) * Nout + c;
      if (K == 0) continue;
      const float* xp = X + (size_t)r0 * ldx;
#pragma unroll 4
      for (int k = 0; k < K; k += 4) {
        const float w0 = __ldg(Wp + (size_t)(k + 0) * Nout);
        const float w1 = __ldg(Wp + (size_t)(k + 1) * Nout);
        const float w2 = __ldg(Wp + (size_t)(k + 2) * Nout);
        const float w3 = __ldg(Wp + (size_t)(k + 3) * Nout);
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          const float4 x = *reinterpret_cast<const float4*>(xp + r * ldx + k);
          acc[r] = fmaf(x.x, w0, acc[r]);
          acc[r] = fmaf(x.y, w1, acc[r]);
          acc[r] = fmaf(x.z, w2, acc[r]);
          acc[r] = fmaf(x.w, w3, acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) Y[(size_t)(r0 + r) * ldy + c] = relu ? fmaxf(acc[r], 0.f) : acc[r];
  }
}

__device__ __forceinline__ void sg_gemm(const float* __restrict__ Xa, int Ka, const float* __restrict__ Xb, int Kb, int ldx,
                                        int RP, const float* __restrict__ W, const float* __restrict__ bias, int Nout,
                                        float* __restrict__ Y, int ldy, bool relu) {
  if (2 * Nout <= SG_THREADS) {
    // narrow outputs (the D = 128 projections): two threads per column, 12 rows of every 24-row chunk each -- all 256
    // threads work instead of half of them
    const int per = SG_THREADS / 2;
    const int half = threadIdx.x / per, t = threadIdx.x - half * per;
    for (int c = t; c < Nout; c += per)
      sg_gemm_rows<SG_RT / 2>(Xa, Ka, Xb, Kb, ldx, RP, W, bias, Nout, Y, ldy, relu, c, half * (SG_RT / 2));
  } else {
    for (int c = threadIdx.x; c < Nout; c += SG_THREADS)
      sg_gemm_rows<SG_RT>(Xa, Ka, Xb, Kb, ldx, RP, W, bias, Nout, Y, ldy, relu, c, 0);
  }
}

struct SgLayerPtrs {
  const float *wq, *bq, *wk, *bk, *wv, *bv, *wm, *bm, *w0, *b0, *w3, *b3;
};

__global__ void __launch_bounds__(SG_THREADS, 1)
superglue_kernel(const float* __restrict__ blob, const __grid_constant__ t2p_superglue_desc d,
                 const float* __restrict__ desc0, const float* __restrict__ desc1, const int64_t* __restrict__ idx0,
                 const int64_t* __restrict__ idx1, int M, int N, int RP, float* __restrict__ outP, int64_t* __restrict__ matches0, int64_t* __restrict__ matches1,
                 float* __restrict__ mscores0, float* __restrict__ mscores1, float* __restrict__ dbg_scores,
                 const int32_t* __restrict__ run_if) {
  extern __shared__ __align__(16) float sg_smem[];
  if (run_if != nullptr && *run_if == 0) return;  // conditional re-run behind the tensor-core kernel (block-uniform)
  const int D = d.dim, R = M + N, dh = D / SG_HEADS;
  const int tid = threadIdx.x, b = blockIdx.x;
  const int maxn = max(M, N);

  float* X = sg_smem;                      // [RP][D] descriptors (rows 0..M-1 side 0, M..R-1 side 1)
  float* QK = X + (size_t)RP * D;          // [2][RP][D]  q | k   (later: hidden [RP][2D])
  float* Vb = QK + 2 * (size_t)RP * D;     // [RP][D]     v       (later: merged message)
  float* MS = Vb + (size_t)RP * D;         // [RP][D]     attention message (later: delta)
  float* PR = MS + (size_t)RP * D;         // [SG_HEADS][R][maxn] attention probabilities
  float* Z = PR + (size_t)SG_HEADS * R * maxn;  // [(M+1)][(N+1)] couplings, then u [M+1], v [N+1]
  float* U = Z + (size_t)(M + 1) * (N + 1);
  float* Vv = U + (M + 1);
  int* I0 = reinterpret_cast<int*>(Vv + (N + 1));  // [M] argmax per row, [N] argmax per column
  int* I1 = I0 + M;

  // gather variant: sample b reads block idx0[b] of a resident table of object encodings and block idx1[b] of the hint encodings
  const size_t b0 = idx0 ? (size_t)idx0[b] : (size_t)b, b1 = idx1 ? (size_t)idx1[b] : (size_t)b;
  for (int t = tid; t < RP * D; t += SG_THREADS) {
    const int r = t / D, c = t - r * D;
    float v = 0.f;
    if (r < M) v = __ldg(desc0 + (b0 * M + r) * D + c);
    else if (r < R) v = __ldg(desc1 + (b1 * N + (r - M)) * D + c);
    X[t] = v;
  }
  __syncthreads();

  float* Q = QK;
  float* Kb = QK + (size_t)RP * D;
  const float inv_sqrt_dh = 1.f / sqrtf((float)dh);

  for (int L = 0; L < d.num_gnn_layers; ++L) {
    const bool cross = d.is_cross[L] != 0;
    sg_gemm(X, D, nullptr, 0, D, RP, blob + d.q[L].w_off, blob + d.q[L].b_off, D, Q, D, false);
    sg_gemm(X, D, nullptr, 0, D, RP, blob + d.k[L].w_off, blob + d.k[L].b_off, D, Kb, D, false);
    sg_gemm(X, D, nullptr, 0, D, RP, blob + d.v[L].w_off, blob + d.v[L].b_off, D, Vb, D, false);
    __syncthreads();
    // attention scores: query row i attends to the rows of its source set
    for (int t = tid; t < SG_HEADS * R * maxn; t += SG_THREADS) {
      const int h = t / (R * maxn), rem = t - h * R * maxn;
      const int i = rem / maxn, j = rem - i * maxn;
      const bool side0 = i < M;
      const int src_is0 = (side0 != cross);  // source set is side 0?
      const int s0 = src_is0 ? 0 : M, sn = src_is0 ? M : N;
      if (j < sn) {
        const float* qp = Q + (size_t)i * D + h;
        const float* kp = Kb + (size_t)(s0 + j) * D + h;
        float acc = 0.f;
        for (int dd = 0; dd < dh; ++dd) acc = fmaf(qp[dd * SG_HEADS], kp[dd * SG_HEADS], acc);
        PR[t] = acc * inv_sqrt_dh;
      }
    }
    __syncthreads();
    for (int t = tid; t < SG_HEADS * R; t += SG_THREADS) {  // softmax over the source
      const int i = t % R;
      const bool side0 = i < M;
      const int sn = (side0 != cross) ? M : N;
      float* p = PR + (size_t)t * maxn;
      float mx = p[0];
      for (int j = 1; j < sn; ++j) mx = fmaxf(mx, p[j]);
      float sum = 0.f;
      for (int j = 0; j < sn; ++j) {
        const float e = expf(p[j] - mx);
        p[j] = e;
        sum += e;
      }
      const float inv = 1.f / sum;
      for (int j = 0; j < sn; ++j) p[j] *= inv;
    }
    __syncthreads();
    for (int t = tid; t < R * D; t += SG_THREADS) {  // message = prob . value
      const int i = t / D, c = t - i * D;
      const int h = c % SG_HEADS;
      const bool side0 = i < M;
      const int src_is0 = (side0 != cross);
      const int s0 = src_is0 ? 0 : M, sn = src_is0 ? M : N;
      const float* p = PR + ((size_t)h * R + i) * maxn;
      float acc = 0.f;
      for (int j = 0; j < sn; ++j) acc = fmaf(p[j], Vb[(size_t)(s0 + j) * D + c], acc);
      MS[t] = acc;
    }
    __syncthreads();
    sg_gemm(MS, D, nullptr, 0, D, RP, blob + d.merge[L].w_off, blob + d.merge[L].b_off, D, Vb, D, false);  // merged
    __syncthreads();
    sg_gemm(X, D, Vb, D, D, RP, blob + d.mlp0[L].w_off, blob + d.mlp0[L].b_off, 2 * D, QK, 2 * D, true);   // hidden
    __syncthreads();
    sg_gemm(QK, 2 * D, nullptr, 0, 2 * D, RP, blob + d.mlp3[L].w_off, blob + d.mlp3[L].b_off, D, MS, D, false);  // delta
    __syncthreads();
    for (int t = tid; t < R * D; t += SG_THREADS) X[t] += MS[t];
    __syncthreads();
  }

  // final projection, scores, couplings
  sg_gemm(X, D, nullptr, 0, D, RP, blob + d.final_proj.w_off, blob + d.final_proj.b_off, D, Q, D, false);
  __syncthreads();
  const float inv_sqrt_d = 1.f / sqrtf((float)D);
  const int Mp = M + 1, Np = N + 1;
  for (int t = tid; t < Mp * Np; t += SG_THREADS) {
    const int i = t / Np, j = t - i * Np;
    float v = d.bin_score;
    if (i < M && j < N) {
      const float* a = Q + (size_t)i * D;
      const float* c = Q + (size_t)(M + j) * D;
      float acc = 0.f;
      for (int k = 0; k < D; ++k) acc = fmaf(a[k], c[k], acc);
      v = acc * inv_sqrt_d;
      if (dbg_scores) dbg_scores[((size_t)b * M + i) * N + j] = v;
    }
    Z[t] = v;
  }
  for (int t = tid; t < Mp; t += SG_THREADS) U[t] = 0.f;
  for (int t = tid; t < Np; t += SG_THREADS) Vv[t] = 0.f;
  __syncthreads();

  // log-space Sinkhorn (models/superglue.py:149-177) on one warp
  const float norm = -logf((float)(M + N));
  if (tid < 32) {
    const float log_mu_bin = logf((float)N) + norm, log_nu_bin = logf((float)M) + norm;
    for (int it = 0; it < d.sinkhorn_iters; ++it) {
      for (int i = tid; i < Mp; i += 32) {  // u = log_mu - logsumexp_j(Z + v)
        const float* z = Z + i * Np;
        float mx = -INFINITY;
        for (int j = 0; j < Np; ++j) mx = fmaxf(mx, z[j] + Vv[j]);
        float s = 0.f;
        for (int j = 0; j < Np; ++j) s += expf(z[j] + Vv[j] - mx);
        U[i] = (i < M ? norm : log_mu_bin) - (mx + logf(s));
      }
      __syncwarp();
      for (int j = tid; j < Np; j += 32) {  // v = log_nu - logsumexp_i(Z + u)
        float mx = -INFINITY;
        for (int i = 0; i < Mp; ++i) mx = fmaxf(mx, Z[i * Np + j] + U[i]);
        float s = 0.f;
        for (int i = 0; i < Mp; ++i) s += expf(Z[i * Np + j] + U[i] - mx);
        Vv[j] = (j < N ? norm : log_nu_bin) - (mx + logf(s));
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int t = tid; t < Mp * Np; t += SG_THREADS) {
    const int i = t / Np, j = t - i * Np;
    const float lp = Z[t] + U[i] + Vv[j] - norm;
    Z[t] = lp;
    outP[(size_t)b * Mp * Np + t] = expf(lp);
  }
  __syncthreads();
  // mutual nearest neighbours over the real rows / columns (first maximum on ties)
  for (int t = tid; t < M + N; t += SG_THREADS) {
    if (t < M) {
      int best = 0;
      float bv = Z[t * Np];
      for (int j = 1; j < N; ++j) {
        const float v = Z[t * Np + j];
        if (v > bv) { bv = v; best = j; }
      }
      I0[t] = best;
    } else {
      const int j = t - M;
      int best = 0;
      float bv = Z[j];
      for (int i = 1; i < M; ++i) {
        const float v = Z[i * Np + j];
        if (v > bv) { bv = v; best = i; }
      }
      I1[j] = best;
    }
  }
  __syncthreads();
  for (int t = tid; t < M + N; t += SG_THREADS) {
    if (t < M) {
      const int i = t, j = I0[i];
      const bool mutual = I1[j] == i;
      const float ms = mutual ? expf(Z[i * Np + j]) : 0.f;
      const bool valid = mutual && ms > d.match_threshold;
      mscores0[(size_t)b * M + i] = ms;
      matches0[(size_t)b * M + i] = valid ? j : -1;
    } else {
      const int j = t - M, i = I1[j];
      const bool mutual1 = I0[i] == j;
      // mscores1 = where(mutual1, mscores0.gather(1, indices1), 0); valid1 = mutual1 & valid0.gather(1, indices1)
      const bool mutual0_i = I1[I0[i]] == i;
      const float ms0_i = mutual0_i ? expf(Z[i * Np + I0[i]]) : 0.f;
      const bool valid0_i = mutual0_i && ms0_i > d.match_threshold;
      mscores1[(size_t)b * N + j] = mutual1 ? ms0_i : 0.f;
      matches1[(size_t)b * N + j] = (mutual1 && valid0_i) ? i : -1;
    }
  }
}

static size_t sg_smem_bytes(int M, int N, int D, int RP) {
  const int R = M + N, maxn = M > N ? M : N;
  size_t f = 5 * (size_t)RP * D + (size_t)SG_HEADS * R * maxn + (size_t)(M + 1) * (N + 1) + (M + 1) + (N + 1) + M + N;
  return f * sizeof(float);
}

}  // namespace t2p

using namespace t2p;

extern "C" {

size_t t2p_superglue_workspace(int B, int M, int N, int D) {
  (void)B; (void)M; (void)N; (void)D;
  return 256;  // the fp16 range flag of the tensor-core kernel; everything else lives in shared / tensor memory
}

int t2p_superglue_forward(const t2p_weights* w, const t2p_superglue_desc* desc, const float* d_desc0, const float* d_desc1,
                          int B, int M, int N, float* d_P, int64_t* d_matches0, int64_t* d_matches1, float* d_mscores0,
                          float* d_mscores1, float* d_dbg_scores, void* d_ws, size_t ws_bytes, t2p_stream stream) {
  return t2p_superglue_forward_gather(w, desc, d_desc0, nullptr, d_desc1, nullptr, B, M, N, d_P, d_matches0, d_matches1, d_mscores0,
                                      d_mscores1, d_dbg_scores, d_ws, ws_bytes, stream);
}

int t2p_superglue_forward_gather(const t2p_weights* w, const t2p_superglue_desc* desc, const float* d_desc0,
                                 const int64_t* d_idx0, const float* d_desc1, const int64_t* d_idx1, int B, int M, int N,
                                 float* d_P, int64_t* d_matches0, int64_t* d_matches1, float* d_mscores0, float* d_mscores1,
                                 float* d_dbg_scores, void* d_ws, size_t ws_bytes, t2p_stream stream) {
  T2P_REQUIRE(w && desc && d_desc0 && d_desc1 && d_P && d_matches0 && d_matches1 && d_mscores0 && d_mscores1,
              T2P_ERR_INVALID, "superglue: null argument");
  if (B <= 0) return T2P_OK;
  const int D = desc->dim;
  T2P_REQUIRE(M >= 1 && N >= 1, T2P_ERR_INVALID, "superglue: M=%d N=%d must be >= 1", M, N);
  T2P_REQUIRE(D >= 4 && D % 4 == 0, T2P_ERR_INVALID, "superglue: dim=%d must be a positive multiple of 4", D);
  T2P_REQUIRE(desc->num_gnn_layers >= 0 && desc->num_gnn_layers <= T2P_MAX_GNN_LAYERS, T2P_ERR_INVALID,
              "superglue: num_gnn_layers=%d outside [0,%d]", desc->num_gnn_layers, T2P_MAX_GNN_LAYERS);
  for (int L = 0; L < desc->num_gnn_layers; ++L) {
    const t2p_linear_desc* ls[] = {&desc->q[L], &desc->k[L], &desc->v[L], &desc->merge[L], &desc->mlp0[L], &desc->mlp3[L]};
    for (const t2p_linear_desc* l : ls)
      T2P_REQUIRE(lin_ok(w, *l) && l->b_off >= 0, T2P_ERR_INVALID, "superglue: layer %d descriptor outside blob", L);
    T2P_REQUIRE(desc->q[L].k == D && desc->q[L].n == D && desc->k[L].k == D && desc->v[L].k == D && desc->merge[L].k == D &&
                    desc->mlp0[L].k == 2 * D && desc->mlp0[L].n == 2 * D && desc->mlp3[L].k == 2 * D && desc->mlp3[L].n == D,
                T2P_ERR_INVALID, "superglue: layer %d has inconsistent dims", L);
  }
  T2P_REQUIRE(lin_ok(w, desc->final_proj) && desc->final_proj.b_off >= 0 && desc->final_proj.k == D && desc->final_proj.n == D,
              T2P_ERR_INVALID, "superglue: final_proj descriptor invalid");
  const int RP = (M + N + SG_RT - 1) / SG_RT * SG_RT;
  const size_t smem = sg_smem_bytes(M, N, D, RP);
  T2P_REQUIRE(smem <= 227 * 1024, T2P_ERR_UNSUPPORTED, "superglue: M=%d N=%d D=%d need %zu bytes of shared memory", M, N, D, smem);
  cudaStream_t s = as_stream(stream);
  // tensor-core path (D = 128, <= 32 rows per side, tensor-core weight images present, workspace for the range flag): several
  // samples per CTA on tcgen05; the exact-fp32 kernel below then only runs if an activation left the fp16 range
  const int32_t* run_if = nullptr;
  if (superglue_tc_supported(desc, M, N) && d_ws != nullptr && ws_bytes >= sizeof(int32_t) &&
      (size_t)desc->tc_w_off + ((size_t)desc->num_gnn_layers * 20 + 2) * 8192 <= w->n_floats) {
    int32_t* flag = static_cast<int32_t*>(d_ws);
    T2P_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t), s));
    T2P_TRY(launch_superglue_tc(w->d_blob, desc, d_desc0, d_idx0, d_desc1, d_idx1, B, M, N, d_P, d_matches0, d_matches1, d_mscores0,
                                d_mscores1, d_dbg_scores, flag, s));
    run_if = flag;
  }
  T2P_CUDA(cudaFuncSetAttribute(superglue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  superglue_kernel<<<B, SG_THREADS, smem, s>>>(w->d_blob, *desc, d_desc0, d_desc1, d_idx0, d_idx1, M, N, RP, d_P, d_matches0, d_matches1,
                                               d_mscores0, d_mscores1, d_dbg_scores, run_if);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

}  // extern "C"
