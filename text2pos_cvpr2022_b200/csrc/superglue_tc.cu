// (a8-a10) SuperGlue head on the tcgen05 tensor cores (D = 128, 4 heads): models/superglue.py:90-146 (attentional GNN),
// :158-177 (log-space optimal transport), :239-330 (final projection, scores, mutual matching).
//
// One CTA = one M = 128 row tile = floor(128 / (M + N)) samples stacked (5 samples of 16 objects + 6 hints), kept on chip
// through all GNN layers; both sides of a layer share its weights (models/superglue.py:144), so every projection is ONE
// [128 x K] x [K x N] product for all rows of all samples of the tile:
//
//   tensor memory (512 columns x 128 lanes, lane = row):  X [0,128) the descriptors, fp32, resident for the whole kernel;
//                 R1 [128,256) / R2 [256,384) / R3 [384,512) accumulators: q,k,v -> merged (R1) -> hidden (R2,R3) -> delta (R1)
//   shared memory: A operand = fp16 hi/lo split of the current activations, K = 128, 128-byte-swizzled K-major (64 KB);
//                  weight ring: 2 stages x 32 KB (one 64-wide K chunk x 128 output channels, host-packed fp16 hi/lo images
//                  of 2^8.W in CONSUMPTION order, streamed by a producer warp with cp.async.bulk from L2);
//                  K / V rows as fp32 for the attention (66 KB, padded pitch)
//   per product: 3 UMMAs per K step (A_hi.W_hi + A_hi.W_lo + A_lo.W_hi, fp32 accumulation; the dropped lo.lo term is
//                ~2^-22 relative: fp32-grade results, tests: matches identical to the oracle / the reference's golden vectors)
//   CUDA cores:  the 4-head softmax attention over <= 32 source rows (q from TMEM, K / V from shared memory, probabilities in
//                registers; the head-major channel order is a host-side permutation of the q/k/v columns and merge rows),
//                the residual update, and per sample one warp for the scores, the 50 log-space Sinkhorn iterations and the
//                mutual-nearest-neighbour matching.
// Activations that leave the fp16 range raise a flag; the host-enqueued exact-fp32 kernel behind this launch then redoes the
// batch (it returns at once otherwise).
#include <cuda_fp16.h>

#include "kernels.h"
#include "sm100.cuh"

namespace t2p {

using namespace sm100;

constexpr int SGT_D = 128;
constexpr int SGT_ROWS = 128;
constexpr int SGT_HEADS = 4;
static_assert(SGT_HEADS * 32 == 128, "head-major layout: 4 heads of 32 channels");
constexpr int SGT_DH = 32;
constexpr int SGT_CWARPS = 8;
constexpr int SGT_CTHREADS = 32 * SGT_CWARPS;
constexpr int SGT_THREADS = SGT_CTHREADS + 32;     // + the weight producer warp
constexpr int SGT_A_CHUNK = SGT_ROWS * 128;        // 16 KB: 128 rows x 64 fp16
constexpr int SGT_A_PART = 2 * SGT_A_CHUNK;        // hi (or lo) of a K = 128 operand
constexpr int SGT_A_BYTES = 2 * SGT_A_PART;        // 64 KB
constexpr int SGT_W_PART = 128 * 128;              // 16 KB: 128 output channels x 64 fp16
constexpr int SGT_W_STAGE = 2 * SGT_W_PART;        // hi + lo
constexpr int SGT_W_STAGES = 3;
constexpr int SGT_KV_PITCH = SGT_D + 4;            // floats per K / V row: 16-byte aligned rows (float4 reads), rows 4 banks apart
constexpr int SGT_KV_BYTES = SGT_ROWS * SGT_KV_PITCH * 4;
constexpr int SGT_STAGES_PER_LAYER = 20;           // q 2, k 2, v 2, merge 2, mlp0 8, mlp3 4
constexpr int SGT_BIAS_PER_LAYER = 4 * 128 + 256 + 128;
constexpr float SGT_UNSCALE = 1.f / 256.f;
constexpr float SGT_AMAX = 60000.f;
constexpr uint32_t SGT_X = 0, SGT_R1 = 128, SGT_R2 = 256, SGT_R3 = 384;

struct SgtBars {
  uint64_t full[SGT_W_STAGES], empty[SGT_W_STAGES];
  uint64_t mma_done;
  uint32_t tmem_slot;
};

__device__ __forceinline__ void sgt_umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void sgt_bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void sgt_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sgt_csync() { asm volatile("bar.sync 1, %0;" ::"n"(SGT_CTHREADS) : "memory"); }
__device__ __forceinline__ void sgt_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
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
__device__ __forceinline__ void sgt_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__host__ __device__ constexpr uint32_t sgt_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);  // fp16 x fp16 -> fp32, both K-major
}

// 8 consecutive activations -> fp16 hi / lo units (16 bytes each)
__device__ __forceinline__ void sgt_split8(const float (&a)[8], uint4& hi4, uint4& lo4, float& amax) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float m = fabsf(a[j]);
    amax = m <= amax ? amax : m;  // NaN sticks
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 hh = __floats2half2_rn(a[2 * j], a[2 * j + 1]);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(a[2 * j] - back.x, a[2 * j + 1] - back.y);
    h[j] = *reinterpret_cast<const uint32_t*>(&hh);
    l[j] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi4 = make_uint4(h[0], h[1], h[2], h[3]);
  lo4 = make_uint4(l[0], l[1], l[2], l[3]);
}
// store unit u (0..7) of 64-wide chunk `chunk` of row `row` into the swizzled A operand
__device__ __forceinline__ void sgt_store_unit(uint8_t* A, int chunk, int row, int u, const uint4& hi4, const uint4& lo4) {
  const uint32_t off = (uint32_t)chunk * SGT_A_CHUNK + (uint32_t)row * 128u + (uint32_t)((u ^ (row & 7)) << 4);
  *reinterpret_cast<uint4*>(A + off) = hi4;
  *reinterpret_cast<uint4*>(A + SGT_A_PART + off) = lo4;
}

// A operand <- act(scale * TMEM[row, col0 + 64 half ..+64) + bias): the thread's row, its 64-column half = K chunk `half`
// (not inlined: seven call sites per layer would otherwise bloat the kernel past the instruction cache)
template <bool RELU>
__device__ __noinline__ void sgt_tmem_to_A(uint32_t tmem_row_addr, uint32_t col0, int half, int row, float scale,
                                              const float* __restrict__ bias, uint8_t* A, float& amax) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_row_addr + col0 + 64 * half + 32 * g, v);
    tmem_ld_wait();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float a[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float x = __uint_as_float(v[8 * u + e]) * scale;
        if (bias) x += __ldg(bias + 64 * half + 32 * g + 8 * u + e);
        a[e] = RELU ? fmaxf(x, 0.f) : x;
      }
      uint4 hi4, lo4;
      sgt_split8(a, hi4, lo4, amax);
      sgt_store_unit(A, half, row, 4 * g + u, hi4, lo4);
    }
  }
}

struct SgtIssue {  // state of the MMA issuer (compute thread 0)
  int stage;
  uint32_t phase;
};

__device__ __forceinline__ bool sgt_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// one weight stage against A chunk `a_chunk`: 4 K steps x 3 products of M = 128, N = 128, K = 16.  Called by ALL lanes of warp
// 0 (warp-uniform control flow); one elected lane issues the tcgen05 instructions -- inside a per-thread branch (`if (tid ==
// 0)`) the compiler wraps every UTCHMMA in an elect / branch loop and the issuing thread, not the tensor pipe, paces the
// products.  `done`: also commit to mma_done (by the lane that issued the MMAs it tracks).
__device__ __forceinline__ void sgt_mma_stage(SgtBars* bars, uint32_t a_addr, uint32_t w_addr, SgtIssue& is, uint32_t tmem_dst,
                                              int a_chunk, bool accumulate, bool done) {
  mbar_wait(&bars->full[is.stage], is.phase);
  tc_fence_after_sync();
  if (sgt_elect_one()) {
    const uint32_t idesc = sgt_idesc(SGT_ROWS, 128);
    const uint32_t wa = w_addr + is.stage * SGT_W_STAGE;
#pragma unroll
    for (int prod = 0; prod < 3; ++prod) {  // A_hi.W_hi, A_hi.W_lo, A_lo.W_hi
      const uint64_t a_desc = umma_desc_sw128_kmajor(a_addr + (prod == 2 ? SGT_A_PART : 0) + a_chunk * SGT_A_CHUNK);
      const uint64_t b_desc = umma_desc_sw128_kmajor(wa + (prod == 1 ? SGT_W_PART : 0));
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) sgt_umma(tmem_dst, a_desc + 2 * ks, b_desc + 2 * ks, idesc, accumulate || prod != 0 || ks != 0);
    }
    umma_commit(&bars->empty[is.stage]);
    if (done) umma_commit(&bars->mma_done);
  }
  __syncwarp();
  if (++is.stage == SGT_W_STAGES) { is.stage = 0; is.phase ^= 1; }
}

#ifdef T2P_SGT_TRACE  // tools/make_sgt_trace.py: %globaltimer stamps of CTA 0 / thread 0 at the phase boundaries of layers 4 and 5
__device__ unsigned long long sgt_trace[64];
#define GTR(i)                                                          \
  do {                                                                  \
    if (blockIdx.x == 0 && threadIdx.x == 0 && (i) < 64) {              \
      unsigned long long t_;                                            \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));            \
      sgt_trace[(i)] = t_;                                              \
    }                                                                   \
  } while (0)
#else
#define GTR(i) do {} while (0)
#endif

template <int MS>  // MS >= max(M, N): compile-time bound of the attention source loop (16 or 32)
__global__ void __launch_bounds__(SGT_THREADS, 1)
superglue_tc_kernel(const float* __restrict__ blob, const __grid_constant__ t2p_superglue_desc d,
                    const float* __restrict__ desc0, const float* __restrict__ desc1, const int64_t* __restrict__ idx0,
                    const int64_t* __restrict__ idx1, int B, int M, int N, float* __restrict__ outP,
                    int64_t* __restrict__ matches0, int64_t* __restrict__ matches1, float* __restrict__ mscores0,
                    float* __restrict__ mscores1, float* __restrict__ dbg_scores, int32_t* __restrict__ overflow_flag) {
  extern __shared__ __align__(1024) uint8_t sgt_raw[];
  if ((smem_u32(sgt_raw) & 1023u) != 0u) __trap();
  uint8_t* A = sgt_raw;                                     // 64 KB
  uint8_t* Wst = A + SGT_A_BYTES;                           // 64 KB
  float* KV = reinterpret_cast<float*>(Wst + SGT_W_STAGES * SGT_W_STAGE);  // [128][129] fp32
  SgtBars* bars = reinterpret_cast<SgtBars*>(reinterpret_cast<uint8_t*>(KV) + SGT_KV_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = M + N;
  const int spt = SGT_ROWS / R;                              // samples per tile
  const int b_first = blockIdx.x * spt;
  const int ns = min(spt, B - b_first);                      // samples of this tile
  const int valid_rows = ns * R;
  const int L = d.num_gnn_layers;
  const uint32_t* w_stream = reinterpret_cast<const uint32_t*>(blob + d.tc_w_off);
  const float* b_stream = blob + d.tc_b_off;

  if (tid == 0) {
    for (int s = 0; s < SGT_W_STAGES; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->mma_done, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(&bars->tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == SGT_CWARPS) {
    // ===== weight producer: the host packed the images in consumption order, one 32 KB stage after the other =====
    if (lane == 0) {
      const int total = L * SGT_STAGES_PER_LAYER + 2;
      int stage = 0;
      uint32_t ph = 0;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(w_stream);
      for (int i = 0; i < total; ++i) {
        mbar_wait(&bars->empty[stage], ph ^ 1);
        mbar_expect_tx(&bars->full[stage], (uint32_t)SGT_W_STAGE);
        const uint32_t dst = smem_u32(Wst + stage * SGT_W_STAGE);
        sgt_bulk_load(dst, src + (size_t)i * SGT_W_STAGE, 16384u, &bars->full[stage]);
        sgt_bulk_load(dst + 16384u, src + (size_t)i * SGT_W_STAGE + 16384u, 16384u, &bars->full[stage]);
        if (++stage == SGT_W_STAGES) { stage = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== compute warps: thread = (row = TMEM lane 32 (warp % 4) + lane, half = warp / 4) =====
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);  // this warp's lane quadrant
    const uint32_t a_addr = smem_u32(A), w_addr = smem_u32(Wst);
    SgtIssue is = {0, 0u};
    uint32_t done_phase = 0;
    float amax = 0.f;
    // sample / side of this row
    const int smp = row / R, rl = row - smp * R;
    const bool row_valid = row < valid_rows;
    const bool side0 = rl < M;
    const int b = b_first + smp;

    // X <- descriptors (zero rows beyond the tile's samples)
    {
      const float* src = nullptr;
      if (row_valid) {
        const size_t blk0 = idx0 ? (size_t)idx0[b] : (size_t)b, blk1 = idx1 ? (size_t)idx1[b] : (size_t)b;
        src = side0 ? desc0 + (blk0 * M + rl) * SGT_D : desc1 + (blk1 * N + (rl - M)) * SGT_D;
      }
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 x = src ? __ldg(reinterpret_cast<const float4*>(src + 64 * half + 32 * g) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[4 * j] = __float_as_uint(x.x); v[4 * j + 1] = __float_as_uint(x.y);
          v[4 * j + 2] = __float_as_uint(x.z); v[4 * j + 3] = __float_as_uint(x.w);
        }
        sgt_tmem_st32(trow + SGT_X + 64 * half + 32 * g, v);
      }
      sgt_tmem_st_wait();
    }

    auto wait_done = [&]() {
      mbar_wait(&bars->mma_done, done_phase);
      done_phase ^= 1;
      tc_fence_after_sync();
    };
    auto publish_A = [&]() {  // A operand written by this thread -> visible to the tensor core; TMEM reads retired
      sgt_fence_async();
      tc_fence_before_sync();
      sgt_csync();
    };

    const float inv_sqrt_dh = 1.f / sqrtf((float)SGT_DH);
    for (int layer = 0; layer < L; ++layer) {
      const int tr0 = (layer == 4 || layer == 5) ? (layer - 4) * 16 : 64;  // trace build only
      (void)tr0;
      GTR(tr0 + 0);
      const float* bl = b_stream + (size_t)layer * SGT_BIAS_PER_LAYER;
      const float *bq = bl, *bk = bl + 128, *bv = bl + 256, *bm = bl + 384, *b0 = bl + 512, *b3 = bl + 768;
      const bool cross = d.is_cross[layer] != 0;
      // ---- q, k, v = X . Wq', Wk', Wv' (head-major columns) ----
      sgt_tmem_to_A<false>(trow, SGT_X, half, row, 1.f, nullptr, A, amax);
      publish_A();
      GTR(tr0 + 1);
      if (warp == 0) {
        tc_fence_after_sync();
#pragma unroll 1
        for (int mat = 0; mat < 3; ++mat)
          for (int kc = 0; kc < 2; ++kc)
            sgt_mma_stage(bars, a_addr, w_addr, is, tmem_base + SGT_R1 + mat * 128, kc, kc != 0, mat == 2 && kc == 1);
      }
      wait_done();
      GTR(tr0 + 2);
      // ---- attention: K rows -> shared memory, scores + softmax in registers, then V rows, message -> A operand ----
      const bool src_is0 = (side0 != cross);
      const int sn = src_is0 ? M : N;
      const int s0 = row_valid ? smp * R + (src_is0 ? 0 : M) : 0;  // (rows beyond the tile's samples never index past the tile)
      {  // K (+ bias): this thread's 64 columns of its row
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t v[32];
          tmem_ld_32x32(trow + SGT_R2 + 64 * half + 32 * g, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            KV[row * SGT_KV_PITCH + 64 * half + 32 * g + j] = fmaf(__uint_as_float(v[j]), SGT_UNSCALE, __ldg(bk + 64 * half + 32 * g + j));
        }
      }
      sgt_csync();
      float p[2][MS];  // probabilities of this thread's two heads (2 half, 2 half + 1)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * half + hh;
        uint32_t qv[32];
        tmem_ld_32x32(trow + SGT_R1 + h * SGT_DH, qv);
        tmem_ld_wait();
        float q[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) q[j] = fmaf(__uint_as_float(qv[j]), SGT_UNSCALE, __ldg(bq + h * SGT_DH + j));
        // scores against the sn source rows of the sample: fully unrolled over the compile-time bound MS with the row index
        // clamped (no branch, no select chain to place a score in its static register) and four partial sums per score (a
        // single accumulator made every score a 32-deep dependent FMA chain: the attention was 43 % of a layer)
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < MS; ++j) {
          const float4* kr = reinterpret_cast<const float4*>(KV + (s0 + min(j, sn - 1)) * SGT_KV_PITCH + h * SGT_DH);
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 kk = kr[c];
            a0 = fmaf(q[4 * c], kk.x, a0);
            a1 = fmaf(q[4 * c + 1], kk.y, a1);
            a2 = fmaf(q[4 * c + 2], kk.z, a2);
            a3 = fmaf(q[4 * c + 3], kk.w, a3);
          }
          const float acc = ((a0 + a1) + (a2 + a3)) * inv_sqrt_dh;
          p[hh][j] = acc;
          mx = (j < sn) ? fmaxf(mx, acc) : mx;
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < MS; ++j) {
          const float e = (j < sn && row_valid) ? expf(p[hh][j] - mx) : 0.f;
          p[hh][j] = e;
          sum += e;
        }
        const float inv = row_valid ? 1.f / sum : 0.f;
#pragma unroll
        for (int j = 0; j < MS; ++j) p[hh][j] *= inv;
      }
      sgt_csync();  // everybody has read K
      {  // V (+ bias) over K
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t v[32];
          tmem_ld_32x32(trow + SGT_R3 + 64 * half + 32 * g, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            KV[row * SGT_KV_PITCH + 64 * half + 32 * g + j] = fmaf(__uint_as_float(v[j]), SGT_UNSCALE, __ldg(bv + 64 * half + 32 * g + j));
        }
      }
      sgt_csync();
      // message[row, h*32 + c] = sum_j p[j] V[src_j, h*32 + c]  ->  A operand (chunk = half: heads 2 half, 2 half + 1)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * half + hh;
        float a[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) a[c] = 0.f;
#pragma unroll
        for (int j = 0; j < MS; ++j) {  // p[hh][j] = 0 beyond sn and for rows outside the tile's samples
          const float pj = p[hh][j];
          const float4* vr = reinterpret_cast<const float4*>(KV + (s0 + min(j, sn - 1)) * SGT_KV_PITCH + h * SGT_DH);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 vv = vr[c];
            a[4 * c] = fmaf(pj, vv.x, a[4 * c]);
            a[4 * c + 1] = fmaf(pj, vv.y, a[4 * c + 1]);
            a[4 * c + 2] = fmaf(pj, vv.z, a[4 * c + 2]);
            a[4 * c + 3] = fmaf(pj, vv.w, a[4 * c + 3]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float a8[8] = {a[8 * u], a[8 * u + 1], a[8 * u + 2], a[8 * u + 3], a[8 * u + 4], a[8 * u + 5], a[8 * u + 6], a[8 * u + 7]};
          uint4 hi4, lo4;
          sgt_split8(a8, hi4, lo4, amax);
          sgt_store_unit(A, half, row, 4 * hh + u, hi4, lo4);
        }
      }
      publish_A();
      GTR(tr0 + 3);
      // ---- merged = message . Wmerge' -> R1 ----
      if (warp == 0) {
        tc_fence_after_sync();
        for (int kc = 0; kc < 2; ++kc) sgt_mma_stage(bars, a_addr, w_addr, is, tmem_base + SGT_R1, kc, kc != 0, kc == 1);
      }
      wait_done();
      GTR(tr0 + 4);
      // ---- hidden = relu([X | merged] . W0 + b0): K = 256 in two A operands, N = 256 in two 128-column blocks (R2, R3) ----
      sgt_tmem_to_A<false>(trow, SGT_X, half, row, 1.f, nullptr, A, amax);
      publish_A();
      GTR(tr0 + 5);
      if (warp == 0) {
        tc_fence_after_sync();
        for (int nb = 0; nb < 2; ++nb)
          for (int kc = 0; kc < 2; ++kc)
            sgt_mma_stage(bars, a_addr, w_addr, is, tmem_base + SGT_R2 + nb * 128, kc, kc != 0, nb == 1 && kc == 1);
      }
      wait_done();
      GTR(tr0 + 6);
      sgt_tmem_to_A<false>(trow, SGT_R1, half, row, SGT_UNSCALE, bm, A, amax);  // merged + bias
      publish_A();
      GTR(tr0 + 7);
      if (warp == 0) {
        tc_fence_after_sync();
        for (int nb = 0; nb < 2; ++nb)
          for (int kc = 0; kc < 2; ++kc)
            sgt_mma_stage(bars, a_addr, w_addr, is, tmem_base + SGT_R2 + nb * 128, kc, true, nb == 1 && kc == 1);
      }
      wait_done();
      GTR(tr0 + 8);
      // ---- delta = hidden . W3 -> R1 ----
      sgt_tmem_to_A<true>(trow, SGT_R2, half, row, SGT_UNSCALE, b0, A, amax);
      publish_A();
      GTR(tr0 + 9);
      if (warp == 0) {
        tc_fence_after_sync();
        for (int kc = 0; kc < 2; ++kc) sgt_mma_stage(bars, a_addr, w_addr, is, tmem_base + SGT_R1, kc, kc != 0, kc == 1);
      }
      wait_done();
      GTR(tr0 + 10);
      sgt_tmem_to_A<true>(trow, SGT_R3, half, row, SGT_UNSCALE, b0 + 128, A, amax);
      publish_A();
      GTR(tr0 + 11);
      if (warp == 0) {
        tc_fence_after_sync();
        for (int kc = 0; kc < 2; ++kc) sgt_mma_stage(bars, a_addr, w_addr, is, tmem_base + SGT_R1, kc, true, kc == 1);
      }
      wait_done();
      GTR(tr0 + 12);
      // ---- X += delta + b3 ----
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t x[32], dl[32];
        tmem_ld_32x32(trow + SGT_X + 64 * half + 32 * g, x);
        tmem_ld_32x32(trow + SGT_R1 + 64 * half + 32 * g, dl);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v = __uint_as_float(x[j]) + fmaf(__uint_as_float(dl[j]), SGT_UNSCALE, __ldg(b3 + 64 * half + 32 * g + j));
          x[j] = __float_as_uint(row_valid ? v : 0.f);
        }
        sgt_tmem_st32(trow + SGT_X + 64 * half + 32 * g, x);
      }
      sgt_tmem_st_wait();
    }

    GTR(40);
    // ---- final projection -> mdesc rows in shared memory ----
    sgt_tmem_to_A<false>(trow, SGT_X, half, row, 1.f, nullptr, A, amax);
    publish_A();
    if (warp == 0) {
      tc_fence_after_sync();
      for (int kc = 0; kc < 2; ++kc) sgt_mma_stage(bars, a_addr, w_addr, is, tmem_base + SGT_R1, kc, kc != 0, kc == 1);
    }
    wait_done();
    {
      const float* bf = b_stream + (size_t)L * SGT_BIAS_PER_LAYER;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t v[32];
        tmem_ld_32x32(trow + SGT_R1 + 64 * half + 32 * g, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          KV[row * SGT_KV_PITCH + 64 * half + 32 * g + j] = fmaf(__uint_as_float(v[j]), SGT_UNSCALE, __ldg(bf + 64 * half + 32 * g + j));
      }
    }
    if (!(amax <= SGT_AMAX) && overflow_flag) atomicOr(overflow_flag, 1);
    tc_fence_before_sync();
    sgt_csync();

    // ---- per sample (one warp each): scores, log-space Sinkhorn, mutual matching (models/superglue.py:149-177,280-322) ----
    const int Mp = M + 1, Np = N + 1;
    const int per = Mp * Np + Mp + Np + M + N;  // floats of scratch per sample (in the A operand region, free now)
    const float inv_sqrt_d = 1.f / sqrtf((float)SGT_D);
    const float norm = -logf((float)(M + N));
    for (int s = warp; s < ns; s += SGT_CWARPS) {
      float* Z = reinterpret_cast<float*>(A) + (size_t)s * per;
      float* U = Z + Mp * Np;
      float* Vv = U + Mp;
      int* I0 = reinterpret_cast<int*>(Vv + Np);
      int* I1 = I0 + M;
      const int bb = b_first + s;
      const float* md = KV + (size_t)(s * R) * SGT_KV_PITCH;
      for (int t = lane; t < Mp * Np; t += 32) {
        const int i = t / Np, j = t - i * Np;
        float v = d.bin_score;
        if (i < M && j < N) {
          const float* a = md + i * SGT_KV_PITCH;
          const float* c = md + (M + j) * SGT_KV_PITCH;
          float acc = 0.f;
          for (int k = 0; k < SGT_D; ++k) acc = fmaf(a[k], c[k], acc);
          v = acc * inv_sqrt_d;
          if (dbg_scores) dbg_scores[((size_t)bb * M + i) * N + j] = v;
        }
        Z[t] = v;
      }
      for (int t = lane; t < Mp; t += 32) U[t] = 0.f;
      for (int t = lane; t < Np; t += 32) Vv[t] = 0.f;
      __syncwarp();
      const float log_mu_bin = logf((float)N) + norm, log_nu_bin = logf((float)M) + norm;
      for (int it = 0; it < d.sinkhorn_iters; ++it) {
        for (int i = lane; i < Mp; i += 32) {  // u = log_mu - logsumexp_j(Z + v)
          const float* z = Z + i * Np;
          float mx = -INFINITY;
          for (int j = 0; j < Np; ++j) mx = fmaxf(mx, z[j] + Vv[j]);
          float sm = 0.f;
          for (int j = 0; j < Np; ++j) sm += expf(z[j] + Vv[j] - mx);
          U[i] = (i < M ? norm : log_mu_bin) - (mx + logf(sm));
        }
        __syncwarp();
        for (int j = lane; j < Np; j += 32) {  // v = log_nu - logsumexp_i(Z + u)
          float mx = -INFINITY;
          for (int i = 0; i < Mp; ++i) mx = fmaxf(mx, Z[i * Np + j] + U[i]);
          float sm = 0.f;
          for (int i = 0; i < Mp; ++i) sm += expf(Z[i * Np + j] + U[i] - mx);
          Vv[j] = (j < N ? norm : log_nu_bin) - (mx + logf(sm));
        }
        __syncwarp();
      }
      for (int t = lane; t < Mp * Np; t += 32) {
        const int i = t / Np, j = t - i * Np;
        const float lp = Z[t] + U[i] + Vv[j] - norm;
        Z[t] = lp;
        outP[(size_t)bb * Mp * Np + t] = expf(lp);
      }
      __syncwarp();
      for (int t = lane; t < M + N; t += 32) {  // first maximum on ties
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
      __syncwarp();
      for (int t = lane; t < M + N; t += 32) {
        if (t < M) {
          const int i = t, j = I0[i];
          const bool mutual = I1[j] == i;
          const float ms = mutual ? expf(Z[i * Np + j]) : 0.f;
          const bool valid = mutual && ms > d.match_threshold;
          mscores0[(size_t)bb * M + i] = ms;
          matches0[(size_t)bb * M + i] = valid ? j : -1;
        } else {
          const int j = t - M, i = I1[j];
          const bool mutual1 = I0[i] == j;
          const bool mutual0_i = I1[I0[i]] == i;
          const float ms0_i = mutual0_i ? expf(Z[i * Np + I0[i]]) : 0.f;
          const bool valid0_i = mutual0_i && ms0_i > d.match_threshold;
          mscores1[(size_t)bb * N + j] = mutual1 ? ms0_i : 0.f;
          matches1[(size_t)bb * N + j] = (mutual1 && valid0_i) ? i : -1;
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  GTR(41);
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

static size_t sgt_smem_bytes() { return (size_t)SGT_A_BYTES + SGT_W_STAGES * SGT_W_STAGE + SGT_KV_BYTES + sizeof(SgtBars) + 64; }

bool superglue_tc_supported(const t2p_superglue_desc* desc, int M, int N) {
  if (desc->dim != SGT_D || desc->tc_w_off < 0 || desc->tc_b_off < 0) return false;
  if (M < 1 || N < 1 || M > 32 || N > 32 || M + N > SGT_ROWS) return false;
  const int per = (M + 1) * (N + 1) + (M + 1) + (N + 1) + M + N;
  return (size_t)(SGT_ROWS / (M + N)) * per * sizeof(float) <= SGT_A_BYTES;
}

int launch_superglue_tc(const float* blob, const t2p_superglue_desc* desc, const float* d_desc0, const int64_t* d_idx0,
                        const float* d_desc1, const int64_t* d_idx1, int B, int M, int N, float* d_P, int64_t* d_matches0,
                        int64_t* d_matches1, float* d_mscores0, float* d_mscores1, float* d_dbg_scores, int32_t* overflow_flag,
                        cudaStream_t s) {
  const size_t smem = sgt_smem_bytes();
  const int spt = SGT_ROWS / (M + N);
  const int grid = (B + spt - 1) / spt;
  if (M <= 16 && N <= 16) {
    T2P_CUDA(cudaFuncSetAttribute(superglue_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    superglue_tc_kernel<16><<<grid, SGT_THREADS, smem, s>>>(blob, *desc, d_desc0, d_desc1, d_idx0, d_idx1, B, M, N, d_P, d_matches0,
                                                           d_matches1, d_mscores0, d_mscores1, d_dbg_scores, overflow_flag);
  } else {
    T2P_CUDA(cudaFuncSetAttribute(superglue_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    superglue_tc_kernel<32><<<grid, SGT_THREADS, smem, s>>>(blob, *desc, d_desc0, d_desc1, d_idx0, d_idx1, B, M, N, d_P, d_matches0,
                                                           d_matches1, d_mscores0, d_mscores1, d_dbg_scores, overflow_flag);
  }
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

}  // namespace t2p

#ifdef T2P_SGT_TRACE
extern "C" int t2p_debug_sgt_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, t2p::sgt_trace, sizeof(t2p::sgt_trace));
}
#endif
