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
// MUFU-based versions for the register-resident kernel: ex2.approx (rel. error ~2^-22) + a correctly rounded
// reciprocal; absolute error ~1e-7 on outputs in [-1,1], far inside the 1e-4 embedding tolerance, and the limits
// are right (exp overflow -> rcp(inf) = 0).
__device__ __forceinline__ float fast_sigmoid(float x) { return __frcp_rn(1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(-2.f, __frcp_rn(1.f + __expf(2.f * x)), 1.f); }

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
// Cluster of CS = H/32 CTAs per (direction, group of NB sequences); CTA rank r owns hidden units
// [32r, 32r+32) = 128 gate columns.  256 threads; thread (warp w, lane = kp*4 + jj) owns the 4 gates of unit
// j = 4w + jj for the k values {4*(8i + kp) + e : i < H/32, e < 4} (H/8 of the H inputs), i.e. H/2 <= 128
// weights in registers.  One step = H/8 broadcast float4 reads of h per sequence, 4*H*NB/8 FMAs per thread,
// a 3-stage recursive-halving reduce-scatter over the 8 kp lanes (28 shuffles) that leaves lane kp with the
// four complete gate pre-activations of sequence b = kp (sequences 8..NB-1: an xor butterfly, lane kp < NB-8
// takes sequence 8+kp), the cell update in registers, a float4 gather over the 4 jj lanes and one 16-byte
// st.async per peer CTA.  The step hand-off is an mbarrier per h buffer in every CTA: the st.async stores of all
// CS CTAs complete NB*H*4 transaction bytes on it, so no cluster-wide barrier / memory fence sits on the
// critical path.  B200 co-schedules at most 15 clusters of 8 CTAs (ncu launch__cluster_max_active), so NB is
// chosen by the host such that 2*ceil(B/NB) <= 14 whenever possible (NB = 10 for B = 64).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t lstm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// 16-byte store into a peer CTA's shared memory that completes 16 transaction bytes on that CTA's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void lstm_bar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lstm_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void lstm_bar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lstm_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lstm_bar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = lstm_smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();  // a protocol bug must not hang the GPU
  }
}

template <int H, int NB>
__global__ void __launch_bounds__(256, 1)
lstm_reg_kernel(const float* __restrict__ xproj, const float* __restrict__ whh_reg, const int32_t* __restrict__ tokens,
                const int32_t* __restrict__ lengths, int B, int T, int V, float* __restrict__ hfinal) {
  constexpr int CS = H / 32;  // CTAs per cluster
  constexpr int NI = H / 32;  // float4 chunks of h per thread and sequence
  constexpr int NX = NB - 8;  // sequences beyond the 8 handled by the reduce-scatter
  static_assert(NX >= 0 && NX <= 4, "NB in [8,12]");
  extern __shared__ __align__(16) float lstm_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS;
  const int dir = cid & 1, group = cid >> 1;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int kp = lane >> 2, jj = lane & 3;
  const int u0 = rank * 32, j = 4 * w + jj;
  const int b0 = group * NB;

  float* hbuf = lstm_smem;                                          // [2][NB][H]
  uint64_t* bars = reinterpret_cast<uint64_t*>(hbuf + 2 * NB * H);  // [2]
  int* tok = reinterpret_cast<int*>(bars + 2);                      // [NB][T]
  int* len = tok + NB * T;                                          // [NB]
  constexpr uint32_t STEP_BYTES = (uint32_t)NB * H * sizeof(float);

  // register-resident weights: Wr[i][e] = the 4 gates (i,f,g,o) of unit j for k = 4*(8i + kp) + e
  float4 Wr[NI][4];
  {
    const float4* src = reinterpret_cast<const float4*>(whh_reg) + ((size_t)(dir * CS + rank) * NI * 4) * 256 + tid;
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) Wr[i][e] = __ldg(src + (size_t)(i * 4 + e) * 256);
  }
  for (int t = tid; t < 2 * NB * H; t += 256) hbuf[t] = 0.f;
  for (int t = tid; t < NB * T; t += 256) {
    const int b = t / T, tt = t - b * T;
    const int v = (b0 + b < B) ? tokens[(size_t)(b0 + b) * T + tt] : 0;
    tok[t] = (v < 0 || v >= V) ? 0 : v;
  }
  if (tid < NB) len[tid] = (b0 + tid < B) ? min(max(lengths[b0 + tid], 0), T) : 0;
  if (tid == 0) {
    lstm_bar_init(&bars[0], 1);
    lstm_bar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    lstm_bar_expect(&bars[0], STEP_BYTES);  // armed for their first use (steps 2 and 1)
    lstm_bar_expect(&bars[1], STEP_BYTES);
  }
  __syncthreads();
  int max_len = 0;
#pragma unroll
  for (int b = 0; b < NB; ++b) max_len = max(max_len, len[b]);
  const int len1 = len[kp];                         // this lane finalises sequence kp ...
  const int len2 = (kp < NX) ? len[8 + kp] : 0;     // ... and sequence 8 + kp when it exists
  cluster_arrive_release();  // barriers initialised and h buffers zeroed in every CTA before remote stores start
  cluster_wait_acquire();

  const float* xp_base = xproj + (size_t)dir * V * 4 * H + u0 + j;
  auto token_at = [&](int b, int L, int step) -> int {
    if (step >= L) return 0;
    return tok[b * T + (dir ? (L - 1 - step) : step)];
  };
  float xn1[4], xn2[4];
  {
    const float* xp = xp_base + (size_t)token_at(kp, len1, 0) * 4 * H;
#pragma unroll
    for (int g = 0; g < 4; ++g) xn1[g] = __ldg(xp + g * H);
    if (NX > 0) {
      const float* xq = xp_base + (size_t)token_at(kp < NX ? 8 + kp : 0, len2, 0) * 4 * H;
#pragma unroll
      for (int g = 0; g < 4; ++g) xn2[g] = __ldg(xq + g * H);
    }
  }
  float c1 = 0.f, h1 = 0.f, c2 = 0.f, h2 = 0.f;
  const uint32_t hbuf_addr = lstm_smem_u32(hbuf), bars_addr = lstm_smem_u32(bars);

  for (int step = 0; step < max_len; ++step) {
    const int cur = step & 1, nxt = cur ^ 1;
    if (step > 0) {  // h_{step-1} from every CTA of the cluster has landed in hbuf[cur]
      lstm_bar_wait(&bars[cur], (uint32_t)(((step - 1) >> 1) & 1));
      if (tid == 0) lstm_bar_expect(&bars[cur], STEP_BYTES);  // re-arm for step + 2
    }
    const float* hcur = hbuf + (size_t)cur * NB * H + 4 * kp;
    float xg1[4], xg2[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      xg1[g] = xn1[g];
      xg2[g] = xn2[g];
    }
    if (step + 1 < max_len) {  // next step's input projection: L2 latency hidden behind the FMAs below
      const float* xp = xp_base + (size_t)token_at(kp, len1, step + 1) * 4 * H;
#pragma unroll
      for (int g = 0; g < 4; ++g) xn1[g] = __ldg(xp + g * H);
      if (NX > 0) {
        const float* xq = xp_base + (size_t)token_at(kp < NX ? 8 + kp : 0, len2, step + 1) * 4 * H;
#pragma unroll
        for (int g = 0; g < 4; ++g) xn2[g] = __ldg(xq + g * H);
      }
    }
    float acc[NB][4];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[b][g] = 0.f;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
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
    // sequences 0..7: reduce-scatter over the 8 kp lanes, lane kp ends with the gates of sequence kp
    float r1[4][4], r2[2][4], gate1[4], gate2[4];
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
        gate1[g] = keep + __shfl_xor_sync(0xffffffffu, send, 4) + xg1[g];
      }
    }
    // sequences 8..NB-1: xor butterfly (every lane gets every sum), lane kp < NX takes sequence 8 + kp
    if (NX > 0) {
#pragma unroll
      for (int g = 0; g < 4; ++g) gate2[g] = 0.f;
#pragma unroll
      for (int x = 0; x < NX; ++x)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v = acc[8 + x][g];
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          if (kp == x) gate2[g] = v + xg2[g];
        }
    }
    {  // cell updates (sequence kp, and 8 + kp where it exists): branch-free so that the two chains interleave
      const float i1 = fast_sigmoid(gate1[0]), f1 = fast_sigmoid(gate1[1]), g1 = fast_tanh(gate1[2]), o1 = fast_sigmoid(gate1[3]);
      const float cn1 = fmaf(f1, c1, i1 * g1);
      const float hn1 = o1 * fast_tanh(cn1);
      const bool a1 = step < len1;
      c1 = a1 ? cn1 : c1;
      h1 = a1 ? hn1 : h1;
      if (NX > 0) {
        const float i2 = fast_sigmoid(gate2[0]), f2 = fast_sigmoid(gate2[1]), g2 = fast_tanh(gate2[2]), o2 = fast_sigmoid(gate2[3]);
        const float cn2 = fmaf(f2, c2, i2 * g2);
        const float hn2 = o2 * fast_tanh(cn2);
        const bool a2 = step < len2;
        c2 = a2 ? cn2 : c2;
        h2 = a2 ? hn2 : h2;
      }
    }
    if (step + 1 < max_len) {
      // h_t[b][u0 + 4w .. +3] gathered over the 4 jj lanes, then one 16-byte st.async per peer CTA
      const uint32_t dst_bar = bars_addr + (uint32_t)nxt * 8u;
      float4 hv4;
      hv4.x = __shfl_sync(0xffffffffu, h1, (lane & ~3) | 0);
      hv4.y = __shfl_sync(0xffffffffu, h1, (lane & ~3) | 1);
      hv4.z = __shfl_sync(0xffffffffu, h1, (lane & ~3) | 2);
      hv4.w = __shfl_sync(0xffffffffu, h1, (lane & ~3) | 3);
      const uint32_t dst1 = hbuf_addr + (uint32_t)(((nxt * NB + kp) * H + u0 + 4 * w) * sizeof(float));
#pragma unroll
      for (int r = jj; r < CS; r += 4) st_async_v4(map_to_rank(dst1, r), hv4, map_to_rank(dst_bar, r));
      if (NX > 0) {
        float4 hx4;
        hx4.x = __shfl_sync(0xffffffffu, h2, (lane & ~3) | 0);
        hx4.y = __shfl_sync(0xffffffffu, h2, (lane & ~3) | 1);
        hx4.z = __shfl_sync(0xffffffffu, h2, (lane & ~3) | 2);
        hx4.w = __shfl_sync(0xffffffffu, h2, (lane & ~3) | 3);
        if (kp < NX) {
          const uint32_t dst2 = hbuf_addr + (uint32_t)(((nxt * NB + 8 + kp) * H + u0 + 4 * w) * sizeof(float));
#pragma unroll
          for (int r = jj; r < CS; r += 4) st_async_v4(map_to_rank(dst2, r), hx4, map_to_rank(dst_bar, r));
        }
      }
    }
  }
  if (b0 + kp < B) hfinal[((size_t)dir * B + b0 + kp) * H + u0 + j] = h1;
  if (NX > 0 && kp < NX && b0 + 8 + kp < B) hfinal[((size_t)dir * B + b0 + 8 + kp) * H + u0 + j] = h2;
  cluster_arrive_release();  // no CTA exits while a peer may still store into its shared memory
  cluster_wait_acquire();
}

template <int H, int NB>
static int lstm_reg_launch(const float* xproj, const float* whh_reg, const int32_t* tokens, const int32_t* lengths, int B,
                           int T, int V, float* hfinal, cudaStream_t s) {
  auto kern = lstm_reg_kernel<H, NB>;
  const size_t smem = (size_t)2 * NB * H * sizeof(float) + 2 * sizeof(uint64_t) + ((size_t)NB * T + NB) * sizeof(int);
  T2P_REQUIRE(smem <= 200 * 1024, T2P_ERR_UNSUPPORTED, "lstm_encode: T=%d needs %zu bytes of shared memory", T, smem);
  if (smem > 48 * 1024) T2P_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int groups = (B + NB - 1) / NB;
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

// sequences per cluster: the smallest NB in {8,10,12} that needs at most 14 clusters of H/32 CTAs (B200 co-schedules 15
// clusters of 8); 12 beyond that (several waves)
template <int H>
static int lstm_reg_dispatch(const float* xproj, const float* whh_reg, const int32_t* tokens, const int32_t* lengths, int B,
                             int T, int V, float* hfinal, cudaStream_t s) {
  const int max_clusters = (H == 256) ? 14 : 28;
  if (2 * ((B + 7) / 8) <= max_clusters) return lstm_reg_launch<H, 8>(xproj, whh_reg, tokens, lengths, B, T, V, hfinal, s);
  if (2 * ((B + 9) / 10) <= max_clusters) return lstm_reg_launch<H, 10>(xproj, whh_reg, tokens, lengths, B, T, V, hfinal, s);
  return lstm_reg_launch<H, 12>(xproj, whh_reg, tokens, lengths, B, T, V, hfinal, s);
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
  T2P_REQUIRE(desc->path >= 0 && desc->path <= 3, T2P_ERR_INVALID, "lstm_encode: path=%d outside [0,3]", desc->path);
  const bool tc_ok = H == 256 && desc->whh_tc_off >= 0 && desc->xproj4_off >= 0;
  T2P_REQUIRE(desc->path != 3 || tc_ok, T2P_ERR_UNSUPPORTED, "lstm_encode: the tensor-core path needs H == 256 and its packed weights");
  if (tc_ok && (desc->path == 0 || desc->path == 3)) {
    T2P_REQUIRE((size_t)desc->whh_tc_off + (size_t)2 * 8 * 32768 <= w->n_floats &&
                    (size_t)desc->xproj4_off + (size_t)2 * V * 4 * H <= w->n_floats,
                T2P_ERR_INVALID, "lstm_encode: tensor-core weights outside the blob");
    Arena a(d_ws, ws_bytes);
    float* hfinal = a.take<float>((size_t)2 * B * H);
    T2P_REQUIRE(a.ok, T2P_ERR_WORKSPACE, "lstm_encode: workspace %zu < %zu bytes", ws_bytes, a.used);
    T2P_TRY(launch_lstm_tc(wptr(w, desc->xproj4_off), wptr(w, desc->whh_tc_off), d_tokens, d_lengths, B, T, V, hfinal, desc->max_groups, s));
    lstm_finalize_kernel<<<(B + 7) / 8, 256, 0, s>>>(hfinal, B, H, normalize, d_out);
    T2P_LAUNCH_CHECK();
    return T2P_OK;
  }
  // H == 32 would be a "cluster" of one CTA: the st.async hand-off needs a real cluster (compute-sanitizer flags it), so it takes
  // the shared-memory kernel below
  if (desc->whh_reg_off >= 0 && (H == 64 || H == 128 || H == 256) && desc->path != 1) {
    T2P_REQUIRE((size_t)desc->whh_reg_off + (size_t)2 * H * 4 * H <= w->n_floats, T2P_ERR_INVALID,
                "lstm_encode: whh_reg outside the weight blob");
    Arena a(d_ws, ws_bytes);
    float* hfinal = a.take<float>((size_t)2 * B * H);
    T2P_REQUIRE(a.ok, T2P_ERR_WORKSPACE, "lstm_encode: workspace %zu < %zu bytes", ws_bytes, a.used);
    const float* wr = wptr(w, desc->whh_reg_off);
    if (H == 256) T2P_TRY(lstm_reg_dispatch<256>(xproj, wr, d_tokens, d_lengths, B, T, V, hfinal, s));
    else if (H == 128) T2P_TRY(lstm_reg_dispatch<128>(xproj, wr, d_tokens, d_lengths, B, T, V, hfinal, s));
    else T2P_TRY(lstm_reg_dispatch<64>(xproj, wr, d_tokens, d_lengths, B, T, V, hfinal, s));
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
