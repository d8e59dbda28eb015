// (a7) all-pairs cosine scores + top-k over the cell database.
//
// Stage 1 (retrieve_partial_kernel): the DB rows are split evenly over one CTA per SM; every CTA stages
// its rows and the query tile in shared memory, computes the fp32 scores with a 4x4 register tile per
// thread and keeps, per query, a warp-distributed sorted list of its kp = k+6 best rows.
// Stage 2 (retrieve_merge_kernel): one CTA per query merges the per-CTA lists, re-scores the kp finalists
// in float64 (the reference ranks in float64, training/coarse.py:100-103,136) and orders them by
// (score desc, index asc).  The float32 pre-selection is certified like the tensor path's: if the k-th exact score is not above
// (kp-th float32 score + the float32 error bound), the CTA rescans the whole DB in float64.
#include "kernels.h"
#include "topk.cuh"

namespace t2p {

constexpr int RT_QT = 64;     // queries per CTA tile
constexpr int RT_KC = 256;    // channels per shared-memory chunk
constexpr int RT_MAX_KP = 32;

__global__ void __launch_bounds__(512)
retrieve_partial_kernel(const float* __restrict__ q, const float* __restrict__ db, int B, int N, int D, int rows_per_cta,
                        int RG, int kp, float* __restrict__ part_s, int32_t* __restrict__ part_i) {
  extern __shared__ __align__(16) float rt_smem[];
  const int pass_rows = 4 * RG;
  const int kc_pitch = ((min(D, RT_KC) + 3) & ~3) + 4;  // multiple of 4 floats: float4 rows stay 16-byte aligned
  float* Qs = rt_smem;                              // [RT_QT][kc_pitch]
  float* Ds = Qs + RT_QT * kc_pitch;                // [pass_rows][kc_pitch]
  float* St = Ds + pass_rows * kc_pitch;            // [RT_QT][pass_rows + 1]
  float* Ls = St + RT_QT * (pass_rows + 1);         // [RT_QT][RT_MAX_KP] running lists (multi-pass only)
  int32_t* Li = reinterpret_cast<int32_t*>(Ls + RT_QT * RT_MAX_KP);

  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  const int q0 = blockIdx.y * RT_QT;
  const int nq = min(RT_QT, B - q0);
  const int row_begin = blockIdx.x * rows_per_cta;
  const int row_end = min(N, row_begin + rows_per_cta);
  const float NEG_INF = __int_as_float(0xff800000);

  for (int t = tid; t < RT_QT * RT_MAX_KP; t += nthreads) {
    Ls[t] = NEG_INF;
    Li[t] = 0x7fffffff;
  }

  for (int r0 = row_begin; r0 < row_end; r0 += pass_rows) {
    const int nr = min(pass_rows, row_end - r0);
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    for (int k0 = 0; k0 < D; k0 += RT_KC) {
      const int kc = min(RT_KC, D - k0);
      const int kc4 = (kc + 3) >> 2;  // float4 slots per row (tail zero padded)
      __syncthreads();
      // stage queries and DB rows (coalesced along channels; rows beyond nq / nr are zero)
      const bool vec = ((D & 3) == 0) && ((k0 & 3) == 0);
      for (int t = tid; t < RT_QT * kc4; t += nthreads) {
        const int r = t / kc4, c = (t - r * kc4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nq) {
          const float* p = q + (size_t)(q0 + r) * D + k0 + c;
          if (vec && c + 3 < kc) v = __ldg(reinterpret_cast<const float4*>(p));
          else {
            if (c + 0 < kc) v.x = __ldg(p + 0);
            if (c + 1 < kc) v.y = __ldg(p + 1);
            if (c + 2 < kc) v.z = __ldg(p + 2);
            if (c + 3 < kc) v.w = __ldg(p + 3);
          }
        }
        *reinterpret_cast<float4*>(Qs + r * kc_pitch + c) = v;
      }
      for (int t = tid; t < pass_rows * kc4; t += nthreads) {
        const int r = t / kc4, c = (t - r * kc4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nr) {
          const float* p = db + (size_t)(r0 + r) * D + k0 + c;
          if (vec && c + 3 < kc) v = __ldg(reinterpret_cast<const float4*>(p));
          else {
            if (c + 0 < kc) v.x = __ldg(p + 0);
            if (c + 1 < kc) v.y = __ldg(p + 1);
            if (c + 2 < kc) v.z = __ldg(p + 2);
            if (c + 3 < kc) v.w = __ldg(p + 3);
          }
        }
        *reinterpret_cast<float4*>(Ds + r * kc_pitch + c) = v;
      }
      __syncthreads();
      // thread tile: queries {tx + 16a}, rows {ty + RG b}
      const float* qp = Qs + tx * kc_pitch;
      const float* dp = Ds + ty * kc_pitch;
      for (int c = 0; c < kc4 * 4; c += 4) {
        float4 qv[4], dv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) qv[a] = *reinterpret_cast<const float4*>(qp + a * 16 * kc_pitch + c);
#pragma unroll
        for (int b = 0; b < 4; ++b) dv[b] = *reinterpret_cast<const float4*>(dp + b * RG * kc_pitch + c);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            acc[a][b] = fmaf(qv[a].x, dv[b].x, acc[a][b]);
            acc[a][b] = fmaf(qv[a].y, dv[b].y, acc[a][b]);
            acc[a][b] = fmaf(qv[a].z, dv[b].z, acc[a][b]);
            acc[a][b] = fmaf(qv[a].w, dv[b].w, acc[a][b]);
          }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) St[(tx + 16 * a) * (pass_rows + 1) + ty + RG * b] = acc[a][b];
    __syncthreads();

    // per-query selection: one warp per query, lanes scan the pass's rows
    for (int ql = warp; ql < nq; ql += nwarps) {
      WarpTopK<float, int32_t> top;
      top.kp = kp;
      top.s = Ls[ql * RT_MAX_KP + lane];
      top.i = Li[ql * RT_MAX_KP + lane];
      const float* srow = St + ql * (pass_rows + 1);
      for (int r = 0; r < nr; r += 32) {
        const int rr = r + lane;
        const bool valid = rr < nr;
        const float sc = valid ? srow[rr] : NEG_INF;
        top.offer(valid, sc, r0 + rr);
      }
      Ls[ql * RT_MAX_KP + lane] = top.s;
      Li[ql * RT_MAX_KP + lane] = top.i;
    }
  }
  __syncthreads();
  // write this CTA's lists: part[cta][q][slot]
  for (int t = tid; t < nq * kp; t += nthreads) {
    const int ql = t / kp, slot = t - ql * kp;
    const size_t o = ((size_t)blockIdx.x * B + q0 + ql) * kp + slot;
    part_s[o] = Ls[ql * RT_MAX_KP + slot];
    part_i[o] = Li[ql * RT_MAX_KP + slot];
  }
}

// one CTA (4 warps) per query
__global__ void __launch_bounds__(128)
retrieve_merge_kernel(const float* __restrict__ q, const float* __restrict__ db, int B, int N, int D, int G, int kp, int k,
                      int64_t idx_base, const float* __restrict__ part_s, const int32_t* __restrict__ part_i,
                      const float* __restrict__ db_norm2_max, double* __restrict__ out_s, int64_t* __restrict__ out_i,
                      int32_t* __restrict__ stats) {
  __shared__ float ws[4][RT_MAX_KP];
  __shared__ int32_t wi[4][RT_MAX_KP];
  __shared__ int32_t fin_i[RT_MAX_KP];
  __shared__ double fin_d[RT_MAX_KP];
  __shared__ float sh_skp;      // float32 score of the kp-th finalist: no other row has a larger float32 score
  __shared__ int sh_nvalid;     // finalists that are real rows (< kp: every row of the DB is a finalist)
  __shared__ int sh_rescan;
  __shared__ double rs_s[4][32];
  __shared__ int64_t rs_i[4][32];
  const int qi = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float NEG_INF = __int_as_float(0xff800000);

  WarpTopK<float, int32_t> top;
  top.init(kp, NEG_INF, 0x7fffffff);
  const int total = G * kp;
  for (int c0 = warp * 32; c0 < total; c0 += 128) {
    const int c = c0 + lane;
    const bool valid = c < total;
    float sc = NEG_INF;
    int32_t id = 0x7fffffff;
    if (valid) {
      const int g = c / kp, slot = c - g * kp;
      const size_t o = ((size_t)g * B + qi) * kp + slot;
      sc = part_s[o];
      id = part_i[o];
    }
    top.offer(valid && id != 0x7fffffff, sc, id);
  }
  ws[warp][lane] = top.s;
  wi[warp][lane] = top.i;
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < 4; ++w) top.offer(lane < kp && wi[w][lane] != 0x7fffffff, ws[w][lane], wi[w][lane]);
    fin_i[lane] = (lane < kp) ? top.i : 0x7fffffff;
    const unsigned real = __ballot_sync(0xffffffffu, lane < kp && top.i != 0x7fffffff);
    const float last = __shfl_sync(0xffffffffu, top.s, kp - 1);
    if (lane == 0) {
      sh_skp = last;
      sh_nvalid = __popc(real);
    }
  }
  __syncthreads();
  // float64 re-scoring of the finalists (lanes stride the channels; fixed-order tree reduction)
  for (int f = warp; f < kp; f += 4) {
    const int32_t id = fin_i[f];
    double acc = 0.0;
    if (id != 0x7fffffff) {
      const float* qp = q + (size_t)qi * D;
      const float* dp = db + (size_t)id * D;
      for (int c = lane; c < D; c += 32) acc = fma((double)__ldg(qp + c), (double)__ldg(dp + c), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) fin_d[f] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    const double NINF = __longlong_as_double(0xfff0000000000000LL);
    const bool have = lane < kp;
    const int32_t my_i = have ? fin_i[lane] : 0x7fffffff;
    const double my_d = (have && my_i != 0x7fffffff) ? fin_d[lane] : NINF;
    int rank = 0;
    for (int g = 0; g < kp; ++g) {
      const int32_t gi = fin_i[g];
      const double gd = (gi != 0x7fffffff) ? fin_d[g] : NINF;
      const bool gb = gd > my_d || (gd == my_d && (gi < my_i || (gi == my_i && g < lane)));
      rank += gb ? 1 : 0;
    }
    // certification (the float32 pre-selection is only a filter): every row that is not a finalist has a float32 score <= the
    // kp-th finalist's, hence an exact score <= that + eps32 |q| max|d|; the ranking of the finalists stands only if the k-th
    // exact score is strictly above this bound -- otherwise (scores near the top denser than float32 resolves, ties across
    // the cut) the CTA rescans the DB in float64 below
    double qq = 0.0;
    for (int c = lane; c < D; c += 32) {
      const double v = (double)__ldg(q + (size_t)qi * D + c);
      qq = fma(v, v, qq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
    const unsigned kth = __ballot_sync(0xffffffffu, have && rank == k - 1 && my_i != 0x7fffffff);
    const double tk = kth ? __shfl_sync(0xffffffffu, my_d, __ffs(kth) - 1) : NINF;
    bool certified = true;
    if (sh_nvalid >= kp) {  // some row may have been filtered out
      const double eps32 = 1.5 * (double)(D + 8) * 5.9604644775390625e-08;  // float32 dot product: (D + O(1)) 2^-24, with margin
      const double bound = (double)sh_skp + eps32 * sqrt(qq) * sqrt((double)__ldg(db_norm2_max) * (1.0 + 1e-5));
      certified = tk > bound;
    }
    if (lane == 0) {
      sh_rescan = certified ? 0 : 1;
      if (stats) atomicAdd(stats + (certified ? 0 : 1), 1);
    }
    if (certified && have && rank < k) {
      const bool ok = my_i != 0x7fffffff;
      out_s[(size_t)qi * k + rank] = ok ? my_d : NINF;
      out_i[(size_t)qi * k + rank] = ok ? idx_base + (int64_t)my_i : (int64_t)-1;
    }
  }
  __syncthreads();
  if (sh_rescan == 0) return;
  // exact float64 rescan of the whole DB for this query (rare)
  {
    const double NINF = __longlong_as_double(0xfff0000000000000LL);
    const int64_t IMAX = 0x7fffffffffffffffLL;
    WarpTopK<double, int64_t> top;
    top.init(k, NINF, IMAX);
    const float* qp = q + (size_t)qi * D;
    for (int r0 = warp * 32; r0 < N; r0 += 128) {  // each warp: 32 consecutive rows per round, lane l keeps row r0 + l
      double mine = NINF;
      for (int rr = 0; rr < 32; ++rr) {
        const int r = r0 + rr;
        if (r < N) {  // warp-uniform
          const float* dp = db + (size_t)r * D;
          double acc = 0.0;
          for (int c = lane; c < D; c += 32) acc = fma((double)__ldg(qp + c), (double)__ldg(dp + c), acc);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
          if (lane == rr) mine = acc;
        }
      }
      top.offer(r0 + lane < N, mine, (int64_t)(r0 + lane));
    }
    rs_s[warp][lane] = top.s;
    rs_i[warp][lane] = top.i;
    __syncthreads();
    if (warp == 0) {
      for (int w = 1; w < 4; ++w) top.offer(lane < k && rs_i[w][lane] != IMAX, rs_s[w][lane], rs_i[w][lane]);
      if (lane < k) {
        const bool ok = top.i != IMAX;
        out_s[(size_t)qi * k + lane] = ok ? top.s : NINF;
        out_i[(size_t)qi * k + lane] = ok ? idx_base + top.i : (int64_t)-1;
      }
    }
  }
}

// merge of R per-shard lists (float64 scores, int64 global indices): one warp per query, 8 queries per CTA (few fat CTAs: small
// CTAs scattered over every SM get in the way of the cluster launches of the text encoder of the next batch)
constexpr int MERGE_WARPS = 8;
__global__ void __launch_bounds__(32 * MERGE_WARPS)
topk_merge_kernel(const double* __restrict__ scores, const int64_t* __restrict__ idx, int R, int B, int k_in, int k_out,
                  double* __restrict__ out_s, int64_t* __restrict__ out_i) {
  const int qi = blockIdx.x * MERGE_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (qi >= B) return;
  const double NINF = __longlong_as_double(0xfff0000000000000LL);
  const int64_t IMAX = 0x7fffffffffffffffLL;
  WarpTopK<double, int64_t> top;
  top.init(k_out, NINF, IMAX);
  const int total = R * k_in;
  for (int c0 = 0; c0 < total; c0 += 32) {
    const int c = c0 + lane;
    bool valid = c < total;
    double sc = NINF;
    int64_t id = IMAX;
    if (valid) {
      const int r = c / k_in, slot = c - r * k_in;
      const size_t o = ((size_t)r * B + qi) * k_in + slot;
      sc = scores[o];
      id = idx[o];
      valid = id >= 0;
    }
    top.offer(valid, sc, id);
  }
  if (lane < k_out) {
    const bool ok = top.i != IMAX;
    out_s[(size_t)qi * k_out + lane] = ok ? top.s : NINF;
    out_i[(size_t)qi * k_out + lane] = ok ? top.i : (int64_t)-1;
  }
}

struct RetrievePlan {
  int G, rows_per_cta, RG, kp, qtiles;
  size_t smem;
};

static int cached_sm_count() {
  static int sm[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sm[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sm[dev] = v;
  }
  return sm[dev];
}

static RetrievePlan make_plan(int B, int N, int D, int k, int sms) {
  RetrievePlan p;
  p.kp = std::min(RT_MAX_KP, k + 6);
  p.qtiles = (B + RT_QT - 1) / RT_QT;
  p.G = std::max(1, std::min(sms, (N + 7) / 8));
  p.rows_per_cta = (N + p.G - 1) / p.G;
  p.G = (N + p.rows_per_cta - 1) / p.rows_per_cta;
  int pass_rows = std::min(128, (p.rows_per_cta + 7) / 8 * 8);
  const int kc_pitch = ((std::min(D, RT_KC) + 3) & ~3) + 4;
  auto smem_of = [&](int pr) {
    return ((size_t)(RT_QT + pr) * kc_pitch + (size_t)RT_QT * (pr + 1) + 2 * (size_t)RT_QT * RT_MAX_KP) * 4;
  };
  while (pass_rows > 8 && smem_of(pass_rows) > 220 * 1024) pass_rows -= 8;  // 227 KB per CTA on sm_100
  p.RG = pass_rows / 4;  // even -> blockDim multiple of 32
  p.smem = smem_of(pass_rows);
  return p;
}

}  // namespace t2p

using namespace t2p;

extern "C" {

size_t t2p_retrieve_topk_workspace(int B, int N, int D, int k) {
  if (B <= 0 || N <= 0 || D <= 0 || k <= 0) return 0;
  // generic path: G never exceeds the SM count of any sm_100 part we target (<= 160)
  const size_t G = 160;
  const size_t kp = std::min(RT_MAX_KP, k + 6);
  size_t generic = align_up(G * B * kp * sizeof(float), 256) + align_up(G * B * kp * sizeof(int32_t), 256) + 256 /* norm bound */;
  size_t tc = 0;
  for (int sms = 1; sms <= 160; ++sms) {  // the plan depends on the SM count / the CTA cap; size for the worst case
    const TcPlan p = tc_plan(B, N, D, k, sms);
    if (p.ok) tc = std::max(tc, tc_workspace_bytes(p, B));
  }
  return std::max(generic, tc);
}

static int retrieve_generic(const float* d_q, const float* d_db, int B, int N, int D, int k, int64_t idx_base,
                            const float* d_db_norm2_max, double* d_out_scores, int64_t* d_out_idx, int32_t* d_stats, void* d_ws,
                            size_t ws_bytes, cudaStream_t s) {
  T2P_REQUIRE(k >= 1 && k + 6 <= RT_MAX_KP, T2P_ERR_UNSUPPORTED, "retrieve_topk: k=%d outside [1,%d]", k, RT_MAX_KP - 6);
  const int sms = std::min(160, cached_sm_count());
  const RetrievePlan p = make_plan(B, N, D, k, sms);
  Arena a(d_ws, ws_bytes);
  float* part_s = a.take<float>((size_t)p.G * B * p.kp);
  int32_t* part_i = a.take<int32_t>((size_t)p.G * B * p.kp);
  float* norm_slot = a.take<float>(1);
  T2P_REQUIRE(a.ok, T2P_ERR_WORKSPACE, "retrieve_topk: workspace %zu < %zu bytes", ws_bytes, a.used);
  if (d_db_norm2_max == nullptr) {  // the certification bound needs max |d|: one extra pass over the DB
    T2P_TRY(launch_row_norm2_max(d_db, N, D, norm_slot, s));
    d_db_norm2_max = norm_slot;
  }
  T2P_CUDA(cudaFuncSetAttribute(retrieve_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  dim3 grid(p.G, p.qtiles);
  retrieve_partial_kernel<<<grid, 16 * p.RG, p.smem, s>>>(d_q, d_db, B, N, D, p.rows_per_cta, p.RG, p.kp, part_s, part_i);
  T2P_LAUNCH_CHECK();
  retrieve_merge_kernel<<<B, 128, 0, s>>>(d_q, d_db, B, N, D, p.G, p.kp, k, idx_base, part_s, part_i, d_db_norm2_max, d_out_scores,
                                          d_out_idx, d_stats);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

int t2p_retrieve_topk_ex(const float* d_q, const float* d_db, int B, int N, int D, int k, int64_t idx_base,
                         const float* d_db_norm2_max, int flags, double* d_out_scores, int64_t* d_out_idx, int32_t* d_stats,
                         void* d_ws, size_t ws_bytes, t2p_stream stream) {
  T2P_REQUIRE(d_q && d_db && d_out_scores && d_out_idx, T2P_ERR_INVALID, "retrieve_topk: null argument");
  T2P_REQUIRE(B > 0 && N > 0 && D > 0, T2P_ERR_INVALID, "retrieve_topk: B=%d N=%d D=%d must be positive", B, N, D);
  cudaStream_t s = as_stream(stream);
  if (!(flags & T2P_RETRIEVE_FORCE_GENERIC)) {
    // a server with several batches in flight cares about SM-time per batch, not latency: T2P_RETRIEVE_MAX_CTAS(n) spreads the
    // scan over at most n CTAs (each streams more DB tiles against its resident query tile; fewer key lists for the select)
    const int cap = (flags >> 8) & 0xff;
    const int sms = std::min(160, cached_sm_count());
    const TcPlan p = tc_plan(B, N, D, k, cap > 0 ? std::min(cap, sms) : sms);
    if (p.ok)
      return launch_retrieve_tc(p, d_q, d_db, B, N, D, k, idx_base, d_db_norm2_max, (flags & T2P_RETRIEVE_FORCE_RESCAN) ? 1 : 0,
                                d_out_scores, d_out_idx, d_stats, d_ws, ws_bytes, s);
  }
  return retrieve_generic(d_q, d_db, B, N, D, k, idx_base, d_db_norm2_max, d_out_scores, d_out_idx, d_stats, d_ws, ws_bytes, s);
}

int t2p_retrieve_topk(const float* d_q, const float* d_db, int B, int N, int D, int k, int64_t idx_base,
                      double* d_out_scores, int64_t* d_out_idx, void* d_ws, size_t ws_bytes, t2p_stream stream) {
  return t2p_retrieve_topk_ex(d_q, d_db, B, N, D, k, idx_base, nullptr, 0, d_out_scores, d_out_idx, nullptr, d_ws, ws_bytes, stream);
}

int t2p_db_row_norm2_max(const float* d_db, int N, int D, float* d_out, t2p_stream stream) {
  T2P_REQUIRE(d_db && d_out && N > 0 && D > 0, T2P_ERR_INVALID, "db_row_norm2_max: bad argument");
  return launch_row_norm2_max(d_db, N, D, d_out, as_stream(stream));
}

int t2p_topk_merge(const double* d_scores, const int64_t* d_idx, int R, int B, int k_in, int k_out, double* d_out_scores,
                   int64_t* d_out_idx, t2p_stream stream) {
  T2P_REQUIRE(d_scores && d_idx && d_out_scores && d_out_idx, T2P_ERR_INVALID, "topk_merge: null argument");
  T2P_REQUIRE(R >= 1 && B >= 1 && k_in >= 1 && k_out >= 1 && k_out <= 32, T2P_ERR_INVALID, "topk_merge: bad sizes");
  topk_merge_kernel<<<(B + MERGE_WARPS - 1) / MERGE_WARPS, 32 * MERGE_WARPS, 0, as_stream(stream)>>>(d_scores, d_idx, R, B, k_in, k_out,
                                                                                                   d_out_scores, d_out_idx);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

}  // extern "C"
