// (a6) text encoder recurrence on the tcgen05 tensor cores (H = 256).
//
// Cluster of 8 CTAs per (direction, group of <= 16 sequences); CTA rank r owns hidden units [32r, 32r+32), i.e. 128
// gate columns (UMMA M), whose slice of W_hh lives in TENSOR MEMORY for the whole sequence as an fp16 hi/lo pair
// (W*2^8 = hi + lo, relative error 2^-22): 2 x 128 TMEM columns (two fp16 per 32-bit column), written once with
// tcgen05.st.  The A operand therefore never touches shared memory again (the SS form re-read 192 KB of weights per
// step, ~1500 cycles of shared-memory bandwidth).  The hidden state lives in shared memory of every CTA as the UMMA B
// operand [16 sequences x 256] fp16 hi/lo (h*2^4 = hi + lo), double buffered, K-major in the NON-swizzled ("interleaved")
// core-matrix layout ordered [source CTA 8][8-sequence half 2][k-unit 4][hi|lo][8 sequences x 16 bytes]: the 32 units x 16
// sequences a CTA produces per step are then ONE contiguous 2 KB slice of every destination's buffer (the 8 units x 8
// sequences of one epilogue warp a 256-byte block of it); the slice is staged in local shared memory and shipped with ONE
// bulk DSMEM copy per destination (cp.async.bulk.shared::cluster.shared::cta, completing bytes on the destination's
// mbarrier; each of the 8 warps of an epilogue set issues one) instead of 32 lanes x 4 st.async of 16 bytes per warp.  Measured on B200 (tools/dsmem_bench.cu, profiles/r02_dsmem_bench.json): the SM-to-SM
// network sustains 21-35 B/clk per SM with bulk copies against 10-12 B/clk with st.async in this traffic pattern, and
// the network was what bound a step.  One step:
//   control warp : wait for h_{t-1} (mbarrier transaction count), fence.proxy.async, one elected lane issues 48
//                  tcgen05.mma kind::f16 (M = 128 gate columns, N = 16 sequences, K = 16; products hi*hi + hi*lo +
//                  lo*hi, fp32 accumulation in TMEM), commit to an mbarrier.  The loop runs warp-uniformly so that the
//                  descriptors stay in uniform registers;
//   8 epilogue warps (warp w: TMEM lane quadrant w%4, sequences 8*(w/4)..+7; TMEM lane = gate column, rows ordered
//                  unit-major so that 4 consecutive lanes hold the gates i,f,g,o of one unit): tcgen05.ld, 4x4 block
//                  transpose over the 4 lanes of a unit (two butterfly stages of predicated selects + shuffles; each
//                  thread then owns one unit for 2 sequences), + input projection (per-token table, L2), cell update in
//                  registers with MUFU-only activations, h -> fp16 hi/lo into the warp's 256-byte block of the 2 KB staging
//                  slice (one per group and buffer parity: it is reused two steps later, when every peer has provably
//                  consumed it), fence.proxy.async, named barrier of the set's 8 warps, then warp w ships the slice to CTA
//                  (rank + w) % 8 with one bulk copy (an elected lane; all operands warp-uniform).
// No cluster barrier, no __syncthreads and no fence sits on the step.  Dropping the lo*lo product and the fp16
// rounding of lo bound the relative error of a product by ~3*2^-22: fp32-grade results (tests: 1e-4 vs the oracle).
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include "kernels.h"
#include "sm100.cuh"

namespace cg = cooperative_groups;

namespace t2p {

using namespace sm100;

constexpr int LTC_H = 256;
constexpr int LTC_CS = 8;        // CTAs per cluster
constexpr int LTC_NS = 16;       // sequences per cluster (UMMA N)
constexpr int LTC_MAXG = 4;       // ping-pong groups of <= 16 sequences per cluster
constexpr int LTC_EPI_SETS = 2;    // epilogue warp sets: set s serves groups s, s + 2 (their latency chains overlap)
constexpr int LTC_EPI_WARPS = 8 * LTC_EPI_SETS;
constexpr int LTC_GPS = LTC_MAXG / LTC_EPI_SETS;       // groups per epilogue set
constexpr int LTC_THREADS = 32 * (LTC_EPI_WARPS + 1);  // warps 0-15: epilogue, warp 16: control
constexpr int LTC_HB_BUF = 16 * 1024;           // one h buffer: 16 K blocks x [k-unit 2][half 2][hi|lo][128 B] = 16 KB
constexpr int LTC_HB_LBO = 256;                 // bytes between the two k-units (core matrices along K) of a K block
constexpr int LTC_HB_SBO = 1024;                // bytes between the two 8-sequence halves (core matrices along N)
constexpr int LTC_STAGE_BYTES = LTC_MAXG * 2 * 2048;  // [group][parity] 2 KB slices = [half 2][k-unit 4][hi|lo][128 B]
constexpr int LTC_TMEM_COLS = 512;              // [0,128) W hi, [128,256) W lo, [256 + 16 g, +16) accumulator of group g
constexpr int LTC_D_COL = 256;
constexpr float LTC_UNSCALE = 1.f / 4096.f;     // 2^-8 (W) * 2^-4 (h)
constexpr float LTC_HSCALE = 16.f;

struct LtcBars {
  uint64_t h_bar[LTC_MAXG][2];  // [group][buffer]
  uint64_t mma_bar[LTC_MAXG];   // [group]
  uint32_t tmem_slot;
};

__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  // fp16 A/B (format 0), fp32 accumulate, both K-major
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ bool ltc_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t ltc_map_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// bulk DSMEM copy: `bytes` from this CTA's shared memory to a (cluster-mapped) address of a peer, completing the bytes on
// the peer's mbarrier
__device__ __forceinline__ void ltc_bulk_copy(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(remote_dst), "r"(local_src), "r"(bytes), "r"(remote_bar)
               : "memory");
}
// K-major operand in the non-swizzled core-matrix layout: core matrix = 8 rows x 16 bytes, contiguous; LBO = byte distance
// between the two core matrices of a K = 16 (fp16) instruction along K, SBO = between 8-row groups along N.
__device__ __forceinline__ uint64_t ltc_desc_interleave_kmajor(uint32_t smem_addr_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr_bytes & 0x3FFFFu) >> 4);
  d |= (uint64_t)(LTC_HB_LBO >> 4) << 16;
  d |= (uint64_t)(LTC_HB_SBO >> 4) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100); swizzle mode 0 = none
  return d;
}
// MUFU-only activations (ex2.approx + rcp.approx, ~2 ulp each, no slow-path branches): absolute error ~2e-7
__device__ __forceinline__ float ltc_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ltc_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ltc_sigmoid(float x) { return ltc_rcp(1.f + ltc_ex2(-1.4426950408889634f * x)); }
__device__ __forceinline__ float ltc_tanh(float x) { return fmaf(-2.f, ltc_rcp(1.f + ltc_ex2(2.885390081777927f * x)), 1.f); }

#ifdef T2P_LSTM_TRACE  // tools/make_lstm_trace.py: %globaltimer stamps of cluster 0 / CTA rank 0, steps 20..23, per group and event
__device__ unsigned long long lstm_trace[4 * LTC_MAXG * 8];
#define LTR(step, g, ev)                                                                     \
  do {                                                                                       \
    if (blockIdx.x == 0 && (step) >= 20 && (step) < 24) {                                    \
      unsigned long long t_;                                                                 \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                 \
      lstm_trace[(((step) - 20) * LTC_MAXG + (g)) * 8 + (ev)] = t_;                          \
    }                                                                                        \
  } while (0)
#else
#define LTR(step, g, ev) do {} while (0)
#endif

__global__ void __launch_bounds__(LTC_THREADS, 1)
lstm_tc_kernel(const float4* __restrict__ xproj4,   // [2][V][H] float4 = gates (i,f,g,o) of (token, unit)
               const uint4* __restrict__ w_img,     // [2][CS][hi|lo][32 k-units][128 rows] x 8 fp16 (2 x 128 x 256 fp16 = 128 KB per slice)
               const int32_t* __restrict__ tokens, const int32_t* __restrict__ lengths, int B, int T, int V,
               float* __restrict__ hfinal, int ns) {
  // ns = sequences of this cluster (<= 64), split into NG = ceil(ns/16) groups of <= 16 that share the W_hh slice in tensor
  // memory and take turns ("ping-pong"): while the h of one group crosses the SM-to-SM network, the MMAs and cell updates
  // of the others run -- a step stays network-bound, but the network is busy all the time, so a batch needs a fraction of
  // the SM-time (at the price of latency).
  // declared 1024-byte aligned (128-byte swizzle atoms): keeps the shared address space visible to the compiler, so the
  // token / staging accesses compile to LDS/STS instead of generic loads
  extern __shared__ __align__(1024) uint8_t ltc_raw[];
  uint8_t* base = ltc_raw;
  if ((smem_u32(base) & 1023u) != 0u) __trap();
  const int NG = (ns + LTC_NS - 1) / LTC_NS;
  int nsg[LTC_MAXG], b0g[LTC_MAXG];  // sequences per group (balanced), first batch row of each group
  uint32_t step_bytes[LTC_MAXG];     // h of a group's real sequences per step: 256 units x (hi + lo) fp16 each
  uint8_t* hb_smem = base;                                          // [NG groups][2 buffers][16 KB]
  uint8_t* stage_smem = hb_smem + (size_t)NG * 2 * LTC_HB_BUF;      // [group][parity][8 warps][hi|lo][8 seqs][8 units] fp16
  LtcBars* bars = reinterpret_cast<LtcBars*>(stage_smem + LTC_STAGE_BYTES);
  int* tok = reinterpret_cast<int*>(bars + 1);                      // [group][NS][T]
  int* len = tok + LTC_MAXG * LTC_NS * T;                           // [group][NS]

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / LTC_CS;
  const int dir = cid & 1, cgroup = cid >> 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u0 = rank * 32;
  {
    int first = cgroup * ns;
#pragma unroll
    for (int g = 0; g < LTC_MAXG; ++g) {  // UMMA columns >= nsg[g] are padding
      nsg[g] = g < NG ? (ns + NG - 1 - g) / NG : 0;
      b0g[g] = first;
      first += nsg[g];
      step_bytes[g] = (uint32_t)((nsg[g] + 7) / 8) * 8192u;  // 8 source CTAs x 1 KB per 8-sequence half in use
    }
  }

  for (int t = tid; t < NG * 2 * LTC_HB_BUF / 16; t += LTC_THREADS) reinterpret_cast<uint4*>(hb_smem)[t] = make_uint4(0, 0, 0, 0);
  for (int t = tid; t < LTC_MAXG * LTC_NS * T; t += LTC_THREADS) {
    const int g = t / (LTC_NS * T), r = t - g * (LTC_NS * T);
    const int b = r / T, tt = r - b * T;
    const int v = (b < nsg[g] && b0g[g] + b < B) ? tokens[(size_t)(b0g[g] + b) * T + tt] : 0;
    tok[t] = (v < 0 || v >= V) ? 0 : v;
  }
  if (tid < LTC_MAXG * LTC_NS) {
    const int g = tid / LTC_NS, b = tid - g * LTC_NS;
    len[tid] = (b < nsg[g] && b0g[g] + b < B) ? min(max(lengths[b0g[g] + b], 0), T) : 0;
  }
  if (tid == 0) {
    for (int i = 0; i < 2 * LTC_MAXG; ++i) mbar_init(&bars->h_bar[0][0] + i, 1);
    for (int g = 0; g < LTC_MAXG; ++g) mbar_init(&bars->mma_bar[g], 1);
    mbar_fence_init();
    for (int g = 0; g < LTC_MAXG; ++g) {
      mbar_expect_tx(&bars->h_bar[g][0], step_bytes[g]);  // armed for their first use (steps 2 and 1)
      mbar_expect_tx(&bars->h_bar[g][1], step_bytes[g]);
    }
  }
  if (warp == LTC_EPI_WARPS) tmem_alloc<LTC_TMEM_COLS>(&bars->tmem_slot);
  fence_proxy_async_smem();  // the zero-filled h buffers (generic stores) will be read by the tensor core
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_slot;
  int max_len[LTC_MAXG];
  int steps = 0;
#pragma unroll
  for (int g = 0; g < LTC_MAXG; ++g) {
    max_len[g] = 0;
#pragma unroll
    for (int b = 0; b < LTC_NS; ++b) max_len[g] = max(max_len[g], len[g * LTC_NS + b]);
    steps = max(steps, max_len[g]);
  }

  if (warp < 8) {
    // W_hh slice -> tensor memory: thread = row m (TMEM lane 32*(warp%4) + lane), warps 0-3 write the hi part, 4-7 the lo
    // part; element k of the row sits in the (k%2) half of 32-bit column k/2.  Coalesced: [k-unit][row] x 16 bytes.
    const int q = warp & 3, part = warp >> 2;
    const uint4* src = w_img + ((size_t)(dir * LTC_CS + rank) * 2 + part) * (32 * 128) + (q * 32 + lane);
#pragma unroll 1
    for (int grp = 0; grp < 4; ++grp) {  // 32 columns = 64 k values = 8 k-units per tcgen05.st
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 x = __ldg(src + (size_t)(grp * 8 + j) * 128);
        v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
      }
      tmem_st_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + part * 128 + grp * 32, v);
    }
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");  // barriers + zeroed buffers visible cluster-wide
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");

  if (warp == LTC_EPI_WARPS) {
    // ===== control warp (warp-uniform loop; one elected lane issues): per step and group wait(h) -> MMAs -> commit.  One
    // barrier per buffer: splitting it per source CTA to start the MMAs of early slices sooner was measured SLOWER (8
    // waits + 8 proxy fences per step cost more than the 48 MMAs they would hide).  Also measured (%globaltimer trace,
    // tools/diag_lstm_trace.py): a step of a group is one dependent chain of ~1.75 us -- MMA issue 0.3, completion + wake-up
    // 0.2, cell update 0.45, staging barrier + copy issue 0.1, SM-to-SM copies until the slowest of the 8 CTAs has delivered
    // 0.6-0.7 -- and the four groups run that chain side by side.  Two control warps (groups {0,2} / {1,3}), polling instead of
    // parked waits, and a dedicated copy warp behind an mbarrier instead of the named barrier all left the chain, hence the
    // step, where it was. =====
    const uint32_t idesc = umma_idesc_f16(128, LTC_NS);
    const uint32_t hb_addr = smem_u32(hb_smem);
    for (int step = 0; step < steps; ++step) {
      const int cur = step & 1;
#pragma unroll
      for (int g = 0; g < LTC_MAXG; ++g) {
        if (step >= max_len[g]) continue;  // warp-uniform
        if (step > 0) {
          mbar_wait(&bars->h_bar[g][cur], (uint32_t)(((step - 1) >> 1) & 1));
          if (lane == 0) mbar_expect_tx(&bars->h_bar[g][cur], step_bytes[g]);  // re-arm for step + 2
          // no proxy fence: h arrives through the async proxy (bulk copies) and the UMMA reads through it too
        }
        if (lane == 0) LTR(step, g, 0);
        tc_fence_after_sync();
        if (ltc_elect_one()) {
          const uint32_t hb = hb_addr + (g * 2 + cur) * LTC_HB_BUF;
          const uint32_t tmem_d = tmem_base + LTC_D_COL + g * LTC_NS;
          bool acc = false;
#pragma unroll
          for (int prod = 0; prod < 3; ++prod) {  // W_hi*h_hi, W_hi*h_lo, W_lo*h_hi
            const uint32_t a_col = tmem_base + (prod == 2 ? 128 : 0);
            const uint64_t b_desc = ltc_desc_interleave_kmajor(hb + (prod == 1 ? 128 : 0));
#pragma unroll
            for (int kb = 0; kb < 16; ++kb) {  // K = 16 fp16 per instruction: 8 TMEM columns of A; B: source CTA kb/2, k-units 2(kb%2), +1
              umma_f16_ts(tmem_d, a_col + kb * 8, b_desc + (uint64_t)(((kb >> 1) * 2048 + (kb & 1) * 512) >> 4), idesc, acc);
              acc = true;
            }
          }
          umma_commit(&bars->mma_bar[g]);
          LTR(step, g, 1);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue warps: TMEM lane m = 32*q + lane  <->  unit 8*q + lane/4, gate lane%4; columns = sequences.  Two sets of
    // 8 warps: set gs serves groups gs and gs + 2.  The epilogue of a group is one long dependent chain (mbarrier wake-up,
    // tcgen05.ld, transposes, six dependent MUFU activations, staging, proxy fence, bulk copies: ~1,500 cycles) -- with one
    // set the chains of the four groups ran back to back and bounded the step; two sets overlap them. =====
    const int q = warp & 3, hf = (warp >> 2) & 1, gs = warp >> 3;
    const int g4 = lane & 3, jj = lane >> 2;
    const bool gb0 = (g4 & 1) != 0, gb1 = (g4 & 2) != 0;
    const int unit = u0 + 8 * q + jj;
    const int s0 = 8 * hf + 2 * g4;  // this thread finalises sequences s0, s0 + 1 of each of its groups
    // this set's groups: gi -> g = gs + 2 gi (selected with static indices: the per-group arrays stay in registers)
    int gq[LTC_GPS], g_nsg[LTC_GPS], g_b0[LTC_GPS], g_maxlen[LTC_GPS];
#pragma unroll
    for (int gi = 0; gi < LTC_GPS; ++gi) {
      gq[gi] = gs + LTC_EPI_SETS * gi;
      g_nsg[gi] = g_b0[gi] = g_maxlen[gi] = 0;
#pragma unroll
      for (int g = 0; g < LTC_MAXG; ++g)
        if (g == gq[gi]) {
          g_nsg[gi] = nsg[g];
          g_b0[gi] = b0g[g];
          g_maxlen[gi] = max_len[g];
        }
    }
    int my_len[LTC_GPS][2] = {};
    float c_state[LTC_GPS][2], h_state[LTC_GPS][2];
    float4 xn[LTC_GPS][2];
    const float4* xp_base = xproj4 + (size_t)dir * V * LTC_H + unit;
    auto token_at = [&](int gi, int e, int step) -> int {
      const int L = my_len[gi][e];
      if (step >= L) return 0;
      return tok[(gq[gi] * LTC_NS + s0 + e) * T + (dir ? (L - 1 - step) : step)];
    };
#pragma unroll
    for (int gi = 0; gi < LTC_GPS; ++gi)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        my_len[gi][e] = len[gq[gi] * LTC_NS + s0 + e];
        c_state[gi][e] = 0.f;
        h_state[gi][e] = 0.f;
        xn[gi][e] = __ldg(xp_base + (size_t)token_at(gi, e, 0) * LTC_H);
      }
    // staging: this warp's [hi|lo][8 seqs][8 units] fp16 = block (half hf, k-unit q) of the CTA's 2 KB slice; warp w of the
    // set ships the whole slice to CTA (rank + w) % 8 -- at any moment the 8 CTAs of the cluster target 8 different peers.
    // Everything that feeds the copy is made provably warp-uniform (shuffle from lane 0): one UBLKCP, no per-lane loop.
    const int lw = __shfl_sync(0xffffffffu, warp & 7, 0);
    const uint32_t dst_rank = (uint32_t)((rank + lw) & (LTC_CS - 1));
    // cluster-mapped addresses of this CTA's slice (group 0, buffer 0) and of h_bar[0][0] in the destination CTA
    const uint32_t rdst = ltc_map_rank(smem_u32(hb_smem) + (uint32_t)rank * 2048u, dst_rank);
    const uint32_t rbar = ltc_map_rank(smem_u32(&bars->h_bar[0][0]), dst_rank);
    const uint32_t tmem_src = tmem_base + ((uint32_t)(q * 32) << 16) + LTC_D_COL + 8 * hf;

    for (int step = 0; step < steps; ++step) {
      const int nxt = (step + 1) & 1;
#pragma unroll
      for (int gi = 0; gi < LTC_GPS; ++gi) {
        if (step >= g_maxlen[gi]) continue;  // warp-uniform
        const int g = gq[gi];
        float4 xg[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) xg[e] = xn[gi][e];
        if (step + 1 < g_maxlen[gi]) {
#pragma unroll
          for (int e = 0; e < 2; ++e) xn[gi][e] = __ldg(xp_base + (size_t)token_at(gi, e, step + 1) * LTC_H);
        }
        mbar_wait(&bars->mma_bar[g], (uint32_t)(step & 1));
        if ((warp & 7) == 0 && lane == 0) LTR(step, g, 2);
        tc_fence_after_sync();
        uint32_t v[8];
        tmem_ld_32x8(tmem_src + g * LTC_NS, v);
        tmem_ld_wait();
        if ((warp & 7) == 0 && lane == 0) LTR(step, g, 3);
        tc_fence_before_sync();
        // 4x4 block transpose (blocks of 2 sequences) over the 4 lanes of a unit as two butterfly stages (xor 2, xor 1);
        // every register index is static and every choice a predicated select, so the warp never diverges.  Lane g4 ends
        // up with, for its sequences: kk0 = gate g4, r0 = gate g4^1, kk1 = gate g4^2, r1 = gate g4^3.
        float ka[4], ra[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t keep = gb1 ? v[4 + i] : v[i];
          const uint32_t send = gb1 ? v[i] : v[4 + i];
          ka[i] = __uint_as_float(keep);
          ra[i] = __uint_as_float(__shfl_xor_sync(0xffffffffu, send, 2));
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float kk0 = gb0 ? ka[2 + e] : ka[e];
          const float sa = gb0 ? ka[e] : ka[2 + e];
          const float kk1 = gb0 ? ra[2 + e] : ra[e];
          const float sb = gb0 ? ra[e] : ra[2 + e];
          const float r0 = __shfl_xor_sync(0xffffffffu, sa, 1);
          const float r1 = __shfl_xor_sync(0xffffffffu, sb, 1);
          const float x = gb0 ? r0 : kk0, y = gb0 ? kk0 : r0, z = gb0 ? r1 : kk1, w = gb0 ? kk1 : r1;
          const float pre_i = gb1 ? z : x, pre_g = gb1 ? x : z, pre_f = gb1 ? w : y, pre_o = gb1 ? y : w;
          // cell update for (unit, sequence s0 + e)
          const float pi = fmaf(pre_i, LTC_UNSCALE, xg[e].x);
          const float pf = fmaf(pre_f, LTC_UNSCALE, xg[e].y);
          const float pg_ = fmaf(pre_g, LTC_UNSCALE, xg[e].z);
          const float po = fmaf(pre_o, LTC_UNSCALE, xg[e].w);
          const float ig = ltc_sigmoid(pi), fg = ltc_sigmoid(pf), gg = ltc_tanh(pg_), og = ltc_sigmoid(po);
          const float cn = fmaf(fg, c_state[gi][e], ig * gg);
          const float hn = og * ltc_tanh(cn);
          const bool active = step < my_len[gi][e];
          c_state[gi][e] = active ? cn : c_state[gi][e];
          h_state[gi][e] = active ? hn : h_state[gi][e];
        }
        if (step + 1 < g_maxlen[gi]) {  // warp-uniform, and uniform over the 8 warps of the set
          // h*2^4 -> fp16 hi/lo, staged as [part][seq-in-half][unit-in-warp]; padding sequences of a half in use carry h = 0;
          // an all-padding second half is neither staged nor shipped (it stays zero in every buffer)
          uint8_t* slice = stage_smem + (size_t)(g * 2 + nxt) * 2048;
          const uint32_t halves = (uint32_t)(g_nsg[gi] + 7) >> 3;
          if ((uint32_t)hf < halves) {
            __half* stage = reinterpret_cast<__half*>(slice + (hf * 4 + q) * 256);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float hs = h_state[gi][e] * LTC_HSCALE;
              const __half hi = __float2half_rn(hs);
              const __half lo = __float2half_rn(hs - __half2float(hi));
              stage[(0 * 8 + 2 * g4 + e) * 8 + jj] = hi;
              stage[(1 * 8 + 2 * g4 + e) * 8 + jj] = lo;
            }
            fence_proxy_async_smem();  // generic-proxy stores -> visible to the bulk copy (async proxy)
          }
          if ((warp & 7) == 0 && lane == 0) LTR(step, g, 4);
          asm volatile("bar.sync %0, 256;" ::"r"(1 + gs) : "memory");  // the 8 warps of this set: the slice is complete
          if ((warp & 7) == 0 && lane == 0) LTR(step, g, 5);
          if (ltc_elect_one())
            ltc_bulk_copy(rdst + (uint32_t)(g * 2 + nxt) * LTC_HB_BUF, smem_u32(slice), halves * 1024u,
                          rbar + (uint32_t)(g * 2 + nxt) * 8u);
          if ((warp & 7) == 0 && lane == 0) LTR(step, g, 6);
        }
      }
    }
#pragma unroll
    for (int gi = 0; gi < LTC_GPS; ++gi)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int b = g_b0[gi] + s0 + e;
        if (s0 + e < g_nsg[gi] && b < B) hfinal[((size_t)dir * B + b) * LTC_H + unit] = h_state[gi][e];
      }
  }
  tc_fence_before_sync();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");  // no CTA exits while a peer may still store into it
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after_sync();
  if (warp == LTC_EPI_WARPS) tmem_dealloc<LTC_TMEM_COLS>(tmem_base);
}

#ifdef T2P_LSTM_TRACE
}  // namespace t2p
extern "C" int t2p_debug_lstm_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, t2p::lstm_trace, sizeof(t2p::lstm_trace));
}
namespace t2p {
#endif

size_t lstm_tc_smem_bytes(int T) {
  return (size_t)LTC_MAXG * 2 * LTC_HB_BUF + LTC_STAGE_BYTES + sizeof(LtcBars) + ((size_t)LTC_MAXG * LTC_NS * T + LTC_MAXG * LTC_NS) * sizeof(int) + 64;
}

int launch_lstm_tc(const float* xproj4, const float* w_img, const int32_t* tokens, const int32_t* lengths, int B, int T, int V,
                   float* hfinal, int max_groups, cudaStream_t s) {
  const size_t smem = lstm_tc_smem_bytes(T);
  T2P_REQUIRE(smem <= 227 * 1024, T2P_ERR_UNSUPPORTED, "lstm_encode: T=%d needs %zu bytes of shared memory", T, smem);
  T2P_CUDA(cudaFuncSetAttribute(lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // The step is bound by the SM-to-SM network (every CTA sends and receives 7/8 KB per sequence of its group and step).
  // Lowest latency (max_groups 0 or 7): spread the batch over as many clusters as B200 co-schedules -- at most 15 clusters
  // of 8 CTAs are resident (launch__cluster_max_active), i.e. 7 clusters x 2 directions, <= 16 sequences each (one group).
  // Throughput (max_groups 1..6): fewer clusters with up to 64 sequences each, run as up to four ping-pong groups that keep
  // the network busy all the time: a fraction of the SM-time per batch, which is what a server with several batches in
  // flight wants.
  const int gmax = (max_groups >= 1 && max_groups <= 7) ? max_groups : 7;
  int groups = B < gmax ? B : gmax;          // clusters per direction
  int ns = (B + groups - 1) / groups;        // sequences per cluster
  if (ns > LTC_MAXG * LTC_NS) ns = LTC_MAXG * LTC_NS;  // large batches: waves of clusters
  groups = (B + ns - 1) / ns;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * 2 * LTC_CS);
  cfg.blockDim = dim3(LTC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = LTC_CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  T2P_CUDA(cudaLaunchKernelEx(&cfg, lstm_tc_kernel, reinterpret_cast<const float4*>(xproj4), reinterpret_cast<const uint4*>(w_img),
                              tokens, lengths, B, T, V, hfinal, ns));
  return T2P_OK;
}

}  // namespace t2p
