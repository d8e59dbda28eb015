// (a6) text encoder: Embedding + packed bidirectional LSTM + mean of final states (+ L2 normalise).
//
// One thread-block CLUSTER per (direction, group of 8 sequences).  The recurrent matrix W_hh^T [H,4H] is
// split by hidden unit over the CTAs of the cluster and stays resident in shared memory for the whole
// sequence (H=256: 8 CTAs x 128 KB); each step every CTA computes the 4 gates of its own hidden units
// for the group's sequences, applies the cell update and broadcasts its slice of h_t into the (double
// buffered) h buffers of all CTAs of the cluster through distributed shared memory, then one cluster
// barrier closes the step.  The input projection is a gather from the per-vocabulary table
// xproj[dir][token] (= emb . W_ih^T + b_ih + b_hh, folded on the host in float64), prefetched one step
// ahead.  Sequences are independent, so clusters never talk to each other.
#include <cooperative_groups.h>

#include "kernels.h"

namespace cg = cooperative_groups;

namespace t2p {

constexpr int LSTM_NB = 8;        // sequences per cluster
constexpr int LSTM_THREADS = 256;

struct LstmSmem {
  float* W;      // [H][NC]
  float* hbuf;   // [2][NB][H]
  float* gates;  // [NB][NC]
  float* cst;    // [NB][HU]
  float* hst;    // [NB][HU]
  int* tok;      // [NB][T]
  int* len;      // [NB]
};

__host__ __device__ inline size_t lstm_smem_floats(int H, int HU, int T) {
  const int NC = 4 * HU;
  return (size_t)H * NC + 2 * (size_t)LSTM_NB * H + (size_t)LSTM_NB * NC + 2 * (size_t)LSTM_NB * HU +
         (size_t)LSTM_NB * T + LSTM_NB;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int RPT>  // rows (sequences) per thread in the gate phase: NB * NC / 256
__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_cluster_kernel(const float* __restrict__ xproj, const float* __restrict__ whh, const int32_t* __restrict__ tokens,
                    const int32_t* __restrict__ lengths, int B, int T, int H, int V, int CS, float* __restrict__ hfinal) {
  extern __shared__ __align__(16) float lstm_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS;
  const int dir = cid & 1, group = cid >> 1;
  const int HU = H / CS, NC = 4 * HU;
  const int u0 = rank * HU;
  const int tid = threadIdx.x;
  const int b0 = group * LSTM_NB;

  LstmSmem sm;
  sm.W = lstm_smem;
  sm.hbuf = sm.W + (size_t)H * NC;
  sm.gates = sm.hbuf + 2 * LSTM_NB * H;
  sm.cst = sm.gates + LSTM_NB * NC;
  sm.hst = sm.cst + LSTM_NB * HU;
  sm.tok = reinterpret_cast<int*>(sm.hst + LSTM_NB * HU);
  sm.len = sm.tok + LSTM_NB * T;

  // resident slice of W_hh^T: W[k][g*HU + j] = whh[dir][k][g*H + u0 + j]
  const float* wsrc = whh + (size_t)dir * H * 4 * H;
  for (int t = tid; t < H * NC; t += LSTM_THREADS) {
    const int k = t / NC, c = t - k * NC;
    const int g = c / HU, j = c - g * HU;
    sm.W[t] = __ldg(wsrc + (size_t)k * 4 * H + g * H + u0 + j);
  }
  for (int t = tid; t < 2 * LSTM_NB * H; t += LSTM_THREADS) sm.hbuf[t] = 0.f;
  for (int t = tid; t < LSTM_NB * HU; t += LSTM_THREADS) {
    sm.cst[t] = 0.f;
    sm.hst[t] = 0.f;
  }
  for (int t = tid; t < LSTM_NB * T; t += LSTM_THREADS) {
    const int b = t / T, tt = t - b * T;
    int v = (b0 + b < B) ? tokens[(size_t)(b0 + b) * T + tt] : 0;
    sm.tok[t] = (v < 0 || v >= V) ? 0 : v;
  }
  if (tid < LSTM_NB) sm.len[tid] = (b0 + tid < B) ? min(max(lengths[b0 + tid], 0), T) : 0;
  __syncthreads();
  int max_len = 0;
#pragma unroll
  for (int b = 0; b < LSTM_NB; ++b) max_len = max(max_len, sm.len[b]);
  cluster.sync();  // every CTA of the cluster is resident and initialised before remote writes start

  const int col = tid % NC, part = tid / NC;
  const int r0 = part * RPT;
  const int gate = col / HU, j_of_col = col - gate * HU;
  const float* xp_base = xproj + (size_t)dir * V * 4 * H + gate * H + u0 + j_of_col;

  auto token_at = [&](int b, int step) -> int {
    const int L = sm.len[b];
    if (step >= L) return 0;
    const int t = dir ? (L - 1 - step) : step;
    return sm.tok[b * T + t];
  };

  float xg[RPT];
#pragma unroll
  for (int r = 0; r < RPT; ++r) xg[r] = __ldg(xp_base + (size_t)token_at(r0 + r, 0) * 4 * H);

  for (int step = 0; step < max_len; ++step) {
    const float* hcur = sm.hbuf + (size_t)(step & 1) * LSTM_NB * H;
    float* hnext_local = sm.hbuf + (size_t)((step + 1) & 1) * LSTM_NB * H;
    float acc[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) acc[r] = xg[r];
    if (step + 1 < max_len) {  // prefetch next step's input projection (L2 latency hidden behind the GEMV)
#pragma unroll
      for (int r = 0; r < RPT; ++r) xg[r] = __ldg(xp_base + (size_t)token_at(r0 + r, step + 1) * 4 * H);
    }
    const float* hp = hcur + (size_t)r0 * H;
    for (int k = 0; k < H; k += 4) {
      const float w0 = sm.W[(k + 0) * NC + col];
      const float w1 = sm.W[(k + 1) * NC + col];
      const float w2 = sm.W[(k + 2) * NC + col];
      const float w3 = sm.W[(k + 3) * NC + col];
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const float4 hv = *reinterpret_cast<const float4*>(hp + r * H + k);
        acc[r] = fmaf(hv.x, w0, acc[r]);
        acc[r] = fmaf(hv.y, w1, acc[r]);
        acc[r] = fmaf(hv.z, w2, acc[r]);
        acc[r] = fmaf(hv.w, w3, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) sm.gates[(r0 + r) * NC + col] = acc[r];
    __syncthreads();

    for (int it = tid; it < LSTM_NB * HU; it += LSTM_THREADS) {
      const int b = it / HU, j = it - b * HU;
      float h = sm.hst[it];
      if (step < sm.len[b]) {
        const float* g = sm.gates + b * NC + j;
        const float ig = sigmoidf_(g[0]);
        const float fg = sigmoidf_(g[HU]);
        const float gg = tanhf(g[2 * HU]);
        const float og = sigmoidf_(g[3 * HU]);
        const float c = fmaf(fg, sm.cst[it], ig * gg);
        h = og * tanhf(c);
        sm.cst[it] = c;
        sm.hst[it] = h;
      }
      const int off = b * H + u0 + j;
      for (int r = 0; r < CS; ++r) {
        float* remote = cluster.map_shared_rank(hnext_local, r);
        remote[off] = h;
      }
    }
    cluster.sync();  // h_{t} complete in every CTA; also orders this step's reads before the next overwrite
  }

  for (int it = tid; it < LSTM_NB * HU; it += LSTM_THREADS) {
    const int b = it / HU, j = it - b * HU;
    if (b0 + b < B) hfinal[((size_t)dir * B + b0 + b) * H + u0 + j] = sm.hst[it];
  }
}

// out[b] = 0.5 * (h_fwd + h_bwd), optionally L2-normalised: one warp per row
__global__ void __launch_bounds__(256)
lstm_finalize_kernel(const float* __restrict__ hfinal, int B, int H, int normalize, float* __restrict__ out) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* f = hfinal + (size_t)b * H;
  const float* r = hfinal + ((size_t)B + b) * H;
  float ss = 0.f;
  for (int c = lane; c < H; c += 32) {
    const float v = 0.5f * (f[c] + r[c]);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float inv = normalize ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : 1.f;
  for (int c = lane; c < H; c += 32) out[(size_t)b * H + c] = 0.5f * (f[c] + r[c]) * inv;
}

// ---------------------------------------------------------------------------------------------------
// Register-resident variant (H in {32,64,128,256}): the recurrent matrix never leaves the register file.
//
// Cluster of CS = H/32 CTAs per (direction, group of 8 sequences); CTA rank r owns hidden units
// [32r, 32r+32) = 128 gate columns.  256 threads; thread (warp w, lane = kp*4 + jj) owns the 4 gates of unit
// j = 4w + jj for the k values {4*(8i + kp) + e : i < H/32, e < 4} (H/8 of the H inputs), i.e. H/2 <= 128
// weights in registers.  One step = H/8 broadcast float4 reads of h per sequence, 4*H FMAs per thread,
// a 3-stage recursive-halving reduce-scatter over the 8 kp lanes (28 shuffles) that leaves lane kp with
// the four complete gate pre-activations of sequence b = kp, the cell update in registers, a float4
// gather over the 4 jj lanes and one 16-byte DSMEM store per peer CTA, then one cluster barrier.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int H>
__global__ void __launch_bounds__(256, 1)
lstm_reg_kernel(const float* __restrict__ xproj, const float* __restrict__ whh_reg, const int32_t* __restrict__ tokens,
                const int32_t* __restrict__ lengths, int B, int T, int V, float* __restrict__ hfinal) {
  constexpr int CS = H / 32;  // CTAs per cluster
  constexpr int NI = H / 32;  // float4 chunks of h per thread and sequence
  extern __shared__ __align__(16) float lstm_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS;
  const int dir = cid & 1, group = cid >> 1;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int kp = lane >> 2, jj = lane & 3;
  const int u0 = rank * 32, j = 4 * w + jj;
  const int b0 = group * LSTM_NB;

  float* hbuf = lstm_smem;                                    // [2][NB][H]
  int* tok = reinterpret_cast<int*>(hbuf + 2 * LSTM_NB * H);  // [NB][T]
  int* len = tok + LSTM_NB * T;                               // [NB]

  // register-resident weights: Wr[i][e] = the 4 gates (i,f,g,o) of unit j for k = 4*(8i + kp) + e
  float4 Wr[NI][4];
  {
    const float4* src = reinterpret_cast<const float4*>(whh_reg) + ((size_t)(dir * CS + rank) * NI * 4) * 256 + tid;
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) Wr[i][e] = __ldg(src + (size_t)(i * 4 + e) * 256);
  }
  for (int t = tid; t < 2 * LSTM_NB * H; t += 256) hbuf[t] = 0.f;
  for (int t = tid; t < LSTM_NB * T; t += 256) {
    const int b = t / T, tt = t - b * T;
    const int v = (b0 + b < B) ? tokens[(size_t)(b0 + b) * T + tt] : 0;
    tok[t] = (v < 0 || v >= V) ? 0 : v;
  }
  if (tid < LSTM_NB) len[tid] = (b0 + tid < B) ? min(max(lengths[b0 + tid], 0), T) : 0;
  __syncthreads();
  int max_len = 0;
#pragma unroll
  for (int b = 0; b < LSTM_NB; ++b) max_len = max(max_len, len[b]);
  const int my_len = len[kp];  // this lane finalises sequence b = kp
  cluster_arrive_release();    // every CTA of the cluster is initialised before remote h writes start
  cluster_wait_acquire();

  const float* xp_base = xproj + (size_t)dir * V * 4 * H + u0 + j;
  auto token_at = [&](int step) -> int {
    if (step >= my_len) return 0;
    return tok[kp * T + (dir ? (my_len - 1 - step) : step)];
  };
  float xn[4];
  {
    const float* xp = xp_base + (size_t)token_at(0) * 4 * H;
#pragma unroll
    for (int g = 0; g < 4; ++g) xn[g] = __ldg(xp + g * H);
  }
  float c_state = 0.f, h_state = 0.f;

  for (int step = 0; step < max_len; ++step) {
    const float* hcur = hbuf + (size_t)(step & 1) * LSTM_NB * H + 4 * kp;
    float* hnext = hbuf + (size_t)((step + 1) & 1) * LSTM_NB * H;
    float xg[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) xg[g] = xn[g];
    if (step + 1 < max_len) {  // next step's input projection: L2 latency hidden behind the FMAs below
      const float* xp = xp_base + (size_t)token_at(step + 1) * 4 * H;
#pragma unroll
      for (int g = 0; g < 4; ++g) xn[g] = __ldg(xp + g * H);
    }
    float acc[LSTM_NB][4];
#pragma unroll
    for (int b = 0; b < LSTM_NB; ++b)
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[b][g] = 0.f;
#pragma unroll
    for (int b = 0; b < LSTM_NB; ++b) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const float4 hv = *reinterpret_cast<const float4*>(hcur + b * H + 32 * i);
        const float he[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[b][0] = fmaf(Wr[i][e].x, he[e], acc[b][0]);
          acc[b][1] = fmaf(Wr[i][e].y, he[e], acc[b][1]);
          acc[b][2] = fmaf(Wr[i][e].z, he[e], acc[b][2]);
          acc[b][3] = fmaf(Wr[i][e].w, he[e], acc[b][3]);
        }
      }
    }
    // reduce-scatter over the 8 kp lanes: lane kp ends with the gates of sequence b = kp
    float r1[4][4], r2[2][4], gate[4];
    {
      const bool up = (kp & 4) != 0;
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float send = up ? acc[t][g] : acc[4 + t][g];
          const float keep = up ? acc[4 + t][g] : acc[t][g];
          r1[t][g] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
      const bool up = (kp & 2) != 0;
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float send = up ? r1[t][g] : r1[2 + t][g];
          const float keep = up ? r1[2 + t][g] : r1[t][g];
          r2[t][g] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
      const bool up = (kp & 1) != 0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float send = up ? r2[0][g] : r2[1][g];
        const float keep = up ? r2[1][g] : r2[0][g];
        gate[g] = keep + __shfl_xor_sync(0xffffffffu, send, 4) + xg[g];
      }
    }
    if (step < my_len) {
      const float ig = sigmoidf_(gate[0]);
      const float fg = sigmoidf_(gate[1]);
      const float gg = tanhf(gate[2]);
      const float og = sigmoidf_(gate[3]);
      c_state = fmaf(fg, c_state, ig * gg);
      h_state = og * tanhf(c_state);
    }
    // h_t[b = kp][u0 + 4w .. +3] gathered over the 4 jj lanes, then one float4 per peer CTA
    float4 hv4;
    hv4.x = __shfl_sync(0xffffffffu, h_state, (lane & ~3) | 0);
    hv4.y = __shfl_sync(0xffffffffu, h_state, (lane & ~3) | 1);
    hv4.z = __shfl_sync(0xffffffffu, h_state, (lane & ~3) | 2);
    hv4.w = __shfl_sync(0xffffffffu, h_state, (lane & ~3) | 3);
    float* dst_local = hnext + kp * H + u0 + 4 * w;
#pragma unroll
    for (int r = jj; r < CS; r += 4) {
      float* remote = cluster.map_shared_rank(dst_local, r);
      *reinterpret_cast<float4*>(remote) = hv4;
    }
    cluster_arrive_release();  // h_t complete in every CTA; also orders this step's reads before the next overwrite
    cluster_wait_acquire();
  }
  if (b0 + kp < B) hfinal[((size_t)dir * B + b0 + kp) * H + u0 + j] = h_state;
}

template <int H>
static int lstm_reg_launch(const float* xproj, const float* whh_reg, const int32_t* tokens, const int32_t* lengths, int B,
                           int T, int V, float* hfinal, cudaStream_t s) {
  auto kern = lstm_reg_kernel<H>;
  const size_t smem = (size_t)2 * LSTM_NB * H * sizeof(float) + ((size_t)LSTM_NB * T + LSTM_NB) * sizeof(int);
  T2P_REQUIRE(smem <= 200 * 1024, T2P_ERR_UNSUPPORTED, "lstm_encode: T=%d needs %zu bytes of shared memory", T, smem);
  if (smem > 48 * 1024) T2P_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int groups = (B + LSTM_NB - 1) / LSTM_NB;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * 2 * (H / 32));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = H / 32;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  T2P_CUDA(cudaLaunchKernelEx(&cfg, kern, xproj, whh_reg, tokens, lengths, B, T, V, hfinal));
  return T2P_OK;
}

struct LstmPlan {
  int CS, HU, RPT;
  size_t smem;
  bool ok;
};

static LstmPlan lstm_plan(int H, int T) {
  LstmPlan p{0, 0, 0, 0, false};
  const int cs_opts[4] = {8, 4, 2, 1};
  for (int i = 0; i < 4; ++i) {
    const int CS = cs_opts[i];
    if (H % CS) continue;
    const int HU = H / CS;
    if (HU != 16 && HU != 32 && HU != 64) continue;
    const size_t smem = lstm_smem_floats(H, HU, T) * sizeof(float);
    if (smem > 200 * 1024) continue;
    p.CS = CS;
    p.HU = HU;
    p.RPT = LSTM_NB * 4 * HU / LSTM_THREADS;
    p.smem = smem;
    p.ok = true;
    return p;
  }
  return p;
}

template <int RPT>
static int lstm_launch(const LstmPlan& p, const float* xproj, const float* whh, const int32_t* tokens, const int32_t* lengths,
                       int B, int T, int H, int V, float* hfinal, cudaStream_t s) {
  auto kern = lstm_cluster_kernel<RPT>;
  T2P_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  const int groups = (B + LSTM_NB - 1) / LSTM_NB;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * 2 * p.CS);
  cfg.blockDim = dim3(LSTM_THREADS);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  T2P_CUDA(cudaLaunchKernelEx(&cfg, kern, xproj, whh, tokens, lengths, B, T, H, V, p.CS, hfinal));
  return T2P_OK;
}

}  // namespace t2p

using namespace t2p;

extern "C" {

size_t t2p_lstm_encode_workspace(int B, int H) {
  if (B <= 0 || H <= 0) return 0;
  return align_up((size_t)2 * B * H * sizeof(float), 256);
}

int t2p_lstm_encode(const t2p_weights* w, const t2p_lstm_desc* desc, const int32_t* d_tokens, const int32_t* d_lengths,
                    int B, int T, int normalize, float* d_out, void* d_ws, size_t ws_bytes, t2p_stream stream) {
  T2P_REQUIRE(w && desc && d_tokens && d_lengths && d_out, T2P_ERR_INVALID, "lstm_encode: null argument");
  if (B <= 0) return T2P_OK;
  const int H = desc->hidden, V = desc->vocab;
  T2P_REQUIRE(T >= 1 && T <= 1024, T2P_ERR_UNSUPPORTED, "lstm_encode: T=%d outside [1,1024]", T);
  T2P_REQUIRE(V >= 1 && H >= 1 && desc->xproj_off >= 0 && desc->whh_off >= 0 &&
                  (size_t)desc->xproj_off + (size_t)2 * V * 4 * H <= w->n_floats &&
                  (size_t)desc->whh_off + (size_t)2 * H * 4 * H <= w->n_floats,
              T2P_ERR_INVALID, "lstm_encode: descriptor outside the weight blob");
  cudaStream_t s = as_stream(stream);
  const float* xproj = wptr(w, desc->xproj_off);
  if (desc->whh_reg_off >= 0 && (H == 32 || H == 64 || H == 128 || H == 256)) {
    T2P_REQUIRE((size_t)desc->whh_reg_off + (size_t)2 * H * 4 * H <= w->n_floats, T2P_ERR_INVALID,
                "lstm_encode: whh_reg outside the weight blob");
    Arena a(d_ws, ws_bytes);
    float* hfinal = a.take<float>((size_t)2 * B * H);
    T2P_REQUIRE(a.ok, T2P_ERR_WORKSPACE, "lstm_encode: workspace %zu < %zu bytes", ws_bytes, a.used);
    const float* wr = wptr(w, desc->whh_reg_off);
    if (H == 256) T2P_TRY(lstm_reg_launch<256>(xproj, wr, d_tokens, d_lengths, B, T, V, hfinal, s));
    else if (H == 128) T2P_TRY(lstm_reg_launch<128>(xproj, wr, d_tokens, d_lengths, B, T, V, hfinal, s));
    else if (H == 64) T2P_TRY(lstm_reg_launch<64>(xproj, wr, d_tokens, d_lengths, B, T, V, hfinal, s));
    else T2P_TRY(lstm_reg_launch<32>(xproj, wr, d_tokens, d_lengths, B, T, V, hfinal, s));
    lstm_finalize_kernel<<<(B + 7) / 8, 256, 0, s>>>(hfinal, B, H, normalize, d_out);
    T2P_LAUNCH_CHECK();
    return T2P_OK;
  }
  const LstmPlan p = lstm_plan(H, T);
  T2P_REQUIRE(p.ok, T2P_ERR_UNSUPPORTED,
              "lstm_encode: hidden=%d not supported (need H = CS*HU with CS in {1,2,4,8}, HU in {16,32,64})", H);
  Arena a(d_ws, ws_bytes);
  float* hfinal = a.take<float>((size_t)2 * B * H);
  T2P_REQUIRE(a.ok, T2P_ERR_WORKSPACE, "lstm_encode: workspace %zu < %zu bytes", ws_bytes, a.used);
  const float* whh = wptr(w, desc->whh_off);
  if (p.RPT == 8) T2P_TRY(lstm_launch<8>(p, xproj, whh, d_tokens, d_lengths, B, T, H, V, hfinal, s));
  else if (p.RPT == 4) T2P_TRY(lstm_launch<4>(p, xproj, whh, d_tokens, d_lengths, B, T, H, V, hfinal, s));
  else T2P_TRY(lstm_launch<2>(p, xproj, whh, d_tokens, d_lengths, B, T, H, V, hfinal, s));
  lstm_finalize_kernel<<<(B + 7) / 8, 256, 0, s>>>(hfinal, B, H, normalize, d_out);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

}  // extern "C"
