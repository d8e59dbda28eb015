// (a6) text encoder recurrence on the tcgen05 tensor cores (H = 256).
//
// Cluster of 8 CTAs per (direction, group of 16 sequences); CTA rank r owns hidden units [32r, 32r+32), i.e. 128
// gate columns, whose slice of W_hh stays in shared memory for the whole sequence as an fp16 hi/lo pair
// (W*2^8 = hi + lo, relative error 2^-22; host-tiled K-major, 128-byte swizzle, loaded with cp.async.bulk).
// The hidden state lives in shared memory of every CTA as the UMMA B operand [16 sequences x 256] fp16 hi/lo
// (h*2^4 = hi + lo), double buffered.  One step:
//   control warp : wait for h_{t-1} (mbarrier transaction count), fence.proxy.async, issue 48 tcgen05.mma
//                  kind::f16 (M = 128 gate columns, N = 16 sequences, K = 16; products hi*hi + hi*lo + lo*hi,
//                  fp32 accumulation in TMEM), commit to an mbarrier;
//   4 epilogue warps (TMEM lane = gate column, rows ordered unit-major so that 4 consecutive lanes hold the gates
//                  i,f,g,o of one unit): tcgen05.ld, 4x4 lane transpose (each thread then owns one unit for 4
//                  sequences), + input projection (per-token table, L2), cell update in registers, h -> fp16 hi/lo,
//                  staged through 512 bytes of shared memory per warp into 16-byte chunks of the swizzled B layout,
//                  one st.async per chunk and peer CTA that completes bytes on the peer's h mbarrier.
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
constexpr int LTC_THREADS = 160; // warps 0-3: epilogue (TMEM quadrants 0-3), warp 4: control
constexpr int LTC_W_BYTES = 2 * 4 * 128 * 128;  // hi|lo x 4 K-chunks x 128 rows x 128 bytes = 128 KB
constexpr int LTC_HB_PART = 4 * LTC_NS * 128;   // one of {hi, lo}: 4 K-chunks x 16 rows x 128 bytes = 8 KB
constexpr int LTC_HB_BUF = 2 * LTC_HB_PART;     // hi + lo
constexpr float LTC_UNSCALE = 1.f / 4096.f;     // 2^-8 (W) * 2^-4 (h)
constexpr float LTC_HSCALE = 16.f;

struct LtcBars {
  uint64_t w_bar;
  uint64_t h_bar[2];
  uint64_t mma_bar;
  uint32_t tmem_slot;
};

__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  // fp16 A/B (format 0), fp32 accumulate, both K-major
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t ltc_map_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void ltc_st_async_v4(uint32_t remote_addr, uint4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ float ltc_sigmoid(float x) { return __frcp_rn(1.f + __expf(-x)); }
__device__ __forceinline__ float ltc_tanh(float x) { return fmaf(-2.f, __frcp_rn(1.f + __expf(2.f * x)), 1.f); }

__global__ void __launch_bounds__(LTC_THREADS, 1)
lstm_tc_kernel(const float4* __restrict__ xproj4,   // [2][V][H] float4 = gates (i,f,g,o) of (token, unit)
               const uint4* __restrict__ w_img,     // [2][CS][LTC_W_BYTES / 16] shared-memory image of the W_hh slice
               const int32_t* __restrict__ tokens, const int32_t* __restrict__ lengths, int B, int T, int V,
               float* __restrict__ hfinal) {
  extern __shared__ uint8_t ltc_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ltc_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* w_smem = base;                                    // [hi: 4 x 16 KB][lo: 4 x 16 KB]
  uint8_t* hb_smem = base + LTC_W_BYTES;                     // [2 buffers][hi|lo][4 chunks][16 x 128 B]
  uint8_t* stage_smem = hb_smem + 2 * LTC_HB_BUF;            // [4 warps][hi|lo][16 seqs][8 units] fp16 = 512 B each
  LtcBars* bars = reinterpret_cast<LtcBars*>(stage_smem + 4 * 512);
  int* tok = reinterpret_cast<int*>(bars + 1);               // [NS][T]
  int* len = tok + LTC_NS * T;                               // [NS]

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / LTC_CS;
  const int dir = cid & 1, group = cid >> 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = group * LTC_NS;
  const int u0 = rank * 32;
  constexpr uint32_t STEP_BYTES = LTC_HB_BUF;  // 16 KB of h (hi + lo, 16 sequences x 256 units) per step

  for (int t = tid; t < 2 * LTC_HB_BUF / 16; t += LTC_THREADS) reinterpret_cast<uint4*>(hb_smem)[t] = make_uint4(0, 0, 0, 0);
  for (int t = tid; t < LTC_NS * T; t += LTC_THREADS) {
    const int b = t / T, tt = t - b * T;
    const int v = (b0 + b < B) ? tokens[(size_t)(b0 + b) * T + tt] : 0;
    tok[t] = (v < 0 || v >= V) ? 0 : v;
  }
  if (tid < LTC_NS) len[tid] = (b0 + tid < B) ? min(max(lengths[b0 + tid], 0), T) : 0;
  if (tid == 0) {
    mbar_init(&bars->w_bar, 1);
    mbar_init(&bars->h_bar[0], 1);
    mbar_init(&bars->h_bar[1], 1);
    mbar_init(&bars->mma_bar, 1);
    mbar_fence_init();
    mbar_expect_tx(&bars->h_bar[0], STEP_BYTES);  // armed for their first use (steps 2 and 1)
    mbar_expect_tx(&bars->h_bar[1], STEP_BYTES);
  }
  if (warp == 4) tmem_alloc<32>(&bars->tmem_slot);
  fence_proxy_async_smem();  // the zero-filled h buffers (generic stores) will be read by the tensor core
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_slot;
  int max_len = 0;
#pragma unroll
  for (int b = 0; b < LTC_NS; ++b) max_len = max(max_len, len[b]);
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");  // barriers + zeroed buffers visible cluster-wide
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");

  if (warp == 4) {
    // ===== control warp: W load, then per step wait(h) -> MMAs -> commit =====
    if (lane == 0) {
      mbar_expect_tx(&bars->w_bar, LTC_W_BYTES);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(w_img) + (size_t)(dir * LTC_CS + rank) * LTC_W_BYTES;
      for (int pc = 0; pc < LTC_W_BYTES / 16384; ++pc) bulk_load(w_smem + pc * 16384, src + pc * 16384, 16384, &bars->w_bar);
      mbar_wait(&bars->w_bar, 0);
      const uint32_t idesc = umma_idesc_f16(128, LTC_NS);
      const uint32_t w_addr = smem_u32(w_smem), hb_addr = smem_u32(hb_smem);
      for (int step = 0; step < max_len; ++step) {
        const int cur = step & 1;
        if (step > 0) {
          mbar_wait(&bars->h_bar[cur], (uint32_t)(((step - 1) >> 1) & 1));
          mbar_expect_tx(&bars->h_bar[cur], STEP_BYTES);  // re-arm for step + 2
        }
        fence_proxy_async_smem();  // h was written through the generic proxy (st.async), the UMMA reads via the async proxy
        tc_fence_after_sync();
        const uint32_t hb = hb_addr + cur * LTC_HB_BUF;
        bool acc = false;
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {  // W_hi*h_hi, W_hi*h_lo, W_lo*h_hi
          const uint32_t wa = w_addr + (prod == 2 ? LTC_W_BYTES / 2 : 0);
          const uint32_t ha = hb + (prod == 1 ? LTC_HB_PART : 0);
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) {
            const uint64_t a_desc = umma_desc_sw128_kmajor(wa + kc * 16384);
            const uint64_t b_desc = umma_desc_sw128_kmajor(ha + kc * (LTC_NS * 128));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {  // K = 16 fp16 = 32 bytes per instruction
              umma_f16_ss(tmem_base, a_desc + 2 * ks, b_desc + 2 * ks, idesc, acc);
              acc = true;
            }
          }
        }
        umma_commit(&bars->mma_bar);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps: TMEM lane m = 32*warp + lane  <->  unit 8*warp + lane/4, gate lane%4 =====
    const int g = lane & 3, jj = lane >> 2;
    const int unit = u0 + 8 * warp + jj;
    int my_len[4];
#pragma unroll
    for (int sq = 0; sq < 4; ++sq) my_len[sq] = len[4 * g + sq];
    float c_state[4] = {0.f, 0.f, 0.f, 0.f}, h_state[4] = {0.f, 0.f, 0.f, 0.f};
    const float4* xp_base = xproj4 + (size_t)dir * V * LTC_H + unit;
    auto token_at = [&](int sq, int step) -> int {
      const int L = my_len[sq];
      if (step >= L) return 0;
      return tok[(4 * g + sq) * T + (dir ? (L - 1 - step) : step)];
    };
    float4 xn[4];
#pragma unroll
    for (int sq = 0; sq < 4; ++sq) xn[sq] = __ldg(xp_base + (size_t)token_at(sq, 0) * LTC_H);
    // staging: this warp's [hi|lo][16 seqs][8 units] fp16; lane l later ships chunk (part = l/16, seq = l%16)
    __half* stage = reinterpret_cast<__half*>(stage_smem + warp * 512);
    const int ship_part = lane >> 4, ship_seq = lane & 15;
    // destination of that chunk inside a B buffer: K index k0 = u0 + 8*warp, chunk k0/64, 16-byte unit (k0%64)/8 ^ (seq&7)
    const int k0 = u0 + 8 * warp;
    const uint32_t ship_off = (uint32_t)(ship_part * LTC_HB_PART + (k0 >> 6) * (LTC_NS * 128) + ship_seq * 128 +
                                         ((((k0 & 63) >> 3) ^ (ship_seq & 7)) << 4));
    const uint32_t hb_addr = smem_u32(hb_smem), hbar_addr = smem_u32(&bars->h_bar[0]);

    for (int step = 0; step < max_len; ++step) {
      const int nxt = (step + 1) & 1;
      float4 xg[4];
#pragma unroll
      for (int sq = 0; sq < 4; ++sq) xg[sq] = xn[sq];
      if (step + 1 < max_len) {
#pragma unroll
        for (int sq = 0; sq < 4; ++sq) xn[sq] = __ldg(xp_base + (size_t)token_at(sq, step + 1) * LTC_H);
      }
      mbar_wait(&bars->mma_bar, (uint32_t)(step & 1));
      tc_fence_after_sync();
      uint32_t v[16];
      tmem_ld_32x16(tmem_base + ((uint32_t)(warp * 32) << 16), v);
      tmem_ld_wait();
      tc_fence_before_sync();
      // 4x4 transpose over the 4 lanes of a unit: afterwards pre[sq][g'] = gate g' of sequence 4g+sq
      float pre[4][4];
#pragma unroll
      for (int sq = 0; sq < 4; ++sq) {
        const float own = __uint_as_float(g == 0 ? v[sq] : g == 1 ? v[4 + sq] : g == 2 ? v[8 + sq] : v[12 + sq]);
#pragma unroll
        for (int gp = 0; gp < 4; ++gp) pre[sq][gp] = own;  // slot gp == g keeps it; the others are overwritten below
      }
#pragma unroll
      for (int r = 1; r < 4; ++r) {
        const int pg = g ^ r;  // partner's gate index = the block of sequences the partner owns
#pragma unroll
        for (int sq = 0; sq < 4; ++sq) {
          const uint32_t send = pg == 0 ? v[sq] : pg == 1 ? v[4 + sq] : pg == 2 ? v[8 + sq] : v[12 + sq];
          const float got = __uint_as_float(__shfl_xor_sync(0xffffffffu, send, r));  // partner's gate pg for my sequences
#pragma unroll
          for (int gp = 0; gp < 4; ++gp)
            if (gp == pg) pre[sq][gp] = got;
        }
      }
      // cell update for (unit, sequences 4g..4g+3)
#pragma unroll
      for (int sq = 0; sq < 4; ++sq) {
        const float pi = fmaf(pre[sq][0], LTC_UNSCALE, xg[sq].x);
        const float pf = fmaf(pre[sq][1], LTC_UNSCALE, xg[sq].y);
        const float pg_ = fmaf(pre[sq][2], LTC_UNSCALE, xg[sq].z);
        const float po = fmaf(pre[sq][3], LTC_UNSCALE, xg[sq].w);
        const float ig = ltc_sigmoid(pi), fg = ltc_sigmoid(pf), gg = ltc_tanh(pg_), og = ltc_sigmoid(po);
        const float cn = fmaf(fg, c_state[sq], ig * gg);
        const float hn = og * ltc_tanh(cn);
        const bool active = step < my_len[sq];
        c_state[sq] = active ? cn : c_state[sq];
        h_state[sq] = active ? hn : h_state[sq];
      }
      if (step + 1 < max_len) {
        // h*2^4 -> fp16 hi/lo, staged as [part][seq][unit-in-warp]
#pragma unroll
        for (int sq = 0; sq < 4; ++sq) {
          const float hs = h_state[sq] * LTC_HSCALE;
          const __half hi = __float2half_rn(hs);
          const __half lo = __float2half_rn(hs - __half2float(hi));
          stage[(0 * LTC_NS + 4 * g + sq) * 8 + jj] = hi;
          stage[(1 * LTC_NS + 4 * g + sq) * 8 + jj] = lo;
        }
        __syncwarp();
        const uint4 chunk = *reinterpret_cast<const uint4*>(stage + (ship_part * LTC_NS + ship_seq) * 8);
        __syncwarp();  // the staging area is rewritten next step
        const uint32_t dst = hb_addr + (uint32_t)nxt * LTC_HB_BUF + ship_off;
        const uint32_t dbar = hbar_addr + (uint32_t)nxt * 8u;
#pragma unroll
        for (int r = 0; r < LTC_CS; ++r) ltc_st_async_v4(ltc_map_rank(dst, r), chunk, ltc_map_rank(dbar, r));
      }
    }
#pragma unroll
    for (int sq = 0; sq < 4; ++sq) {
      const int b = b0 + 4 * g + sq;
      if (b < B) hfinal[((size_t)dir * B + b) * LTC_H + unit] = h_state[sq];
    }
  }
  tc_fence_before_sync();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");  // no CTA exits while a peer may still store into it
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after_sync();
  if (warp == 4) tmem_dealloc<32>(tmem_base);
}

size_t lstm_tc_smem_bytes(int T) {
  return 1024 + (size_t)LTC_W_BYTES + 2 * LTC_HB_BUF + 4 * 512 + sizeof(LtcBars) + ((size_t)LTC_NS * T + LTC_NS) * sizeof(int) + 64;
}

int launch_lstm_tc(const float* xproj4, const float* w_img, const int32_t* tokens, const int32_t* lengths, int B, int T, int V,
                   float* hfinal, cudaStream_t s) {
  const size_t smem = lstm_tc_smem_bytes(T);
  T2P_REQUIRE(smem <= 227 * 1024, T2P_ERR_UNSUPPORTED, "lstm_encode: T=%d needs %zu bytes of shared memory", T, smem);
  T2P_CUDA(cudaFuncSetAttribute(lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int groups = (B + LTC_NS - 1) / LTC_NS;
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
                              tokens, lengths, B, T, V, hfinal));
  return T2P_OK;
}

}  // namespace t2p
