// (a7) all-pairs scores + top-k on the tcgen05 tensor cores: TMA-fed TF32 scan with an in-kernel top-k,
// followed by a certified exact (float64) re-rank.
//
//   scan   (retrieve_scan_tc_kernel): persistent CTAs, one per SM.  The query tile [128 x D] stays resident in
//           shared memory (TMA, 128-byte swizzle); DB row tiles [tile_n x 32 floats] stream through a 4-stage
//           TMA/mbarrier ring; one elected thread issues tcgen05.mma kind::tf32 (M = 128 queries, N = tile_n
//           rows, K = 8 per instruction) into a double-buffered TMEM accumulator; four epilogue warps read their
//           TMEM lane quadrant (lane = query) with tcgen05.ld and keep, per query, the KP best rows of the CTA as
//           sorted 32-bit keys in registers (score bits with the low bits replaced by the local row index --
//           min/max insertion, no divergence).
//   select (retrieve_select_kernel): one CTA per query.  Radix-selects the NC best keys over all CTAs,
//           re-scores those candidates in float64 (the reference ranks float64 dot products of the float32
//           embeddings, training/coarse.py:100-103,136), orders them by (score desc, index asc) and CERTIFIES
//           the result: every row that is not a candidate has a TF32 score <= U, hence a true score
//           <= U + eps*|q|*max|d|; if the k-th exact score is not above that bound the CTA falls back to an exact
//           float64 scan of the whole DB for its query.  The output therefore always equals the float64 ranking;
//           TF32 only decides how fast it is obtained.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "kernels.h"
#include "sm100.cuh"
#include "topk.cuh"

namespace t2p {

using namespace sm100;

constexpr int TC_QM = 128;          // queries per tile (UMMA M)
constexpr int TC_TILE_MAX = 128;    // DB rows per tile (UMMA N), multiple of 16
constexpr int TC_EPI_WARPS = 8;     // at most two per TMEM lane quadrant (they take alternate 16-column chunks of every tile)
constexpr int TC_THREADS = 32 * (2 + TC_EPI_WARPS);  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..: epilogue
constexpr int TC_MAX_TILES_PER_CTA = 16;
constexpr int TC_TMEM_COLS = 512;   // [0, 256): 2 accumulators x 128 columns; [256, 256 + D): the query tile (A operand)
constexpr int TC_Q_COL = 256;
constexpr int TC_EPI_DEFAULT = 4;
constexpr int TC_CAND_CAP = 32;     // candidate keys a lane may buffer between two merges into its sorted list
constexpr int TC_CAND_BYTES = TC_EPI_WARPS * TC_CAND_CAP * 32 * 4;  // epilogue warps x CAP x 32 lanes x u32 = 32 KB
constexpr int SEL_THREADS = 256;
constexpr int SEL_CAND_MAX = 64;
// |tf32 score - exact score| <= TC_EPS * |q| * |d|: both operands lose < 2^-10 relative (13 mantissa bits dropped),
// plus the fp32 accumulation of <= 256 products inside the tensor core.
constexpr double TC_EPS = 2.2e-3;

// ---- order-preserving 32-bit keys ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t score_to_key(float s, uint32_t low_mask, uint32_t local) {
  uint32_t u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (u & ~low_mask) | (low_mask - local);  // lower local index wins among equal (truncated) scores
}
// largest score whose key could be `key`
__device__ __forceinline__ float key_upper_score(uint32_t key, uint32_t low_mask) {
  uint32_t u = key | low_mask;
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(u);
}

// smallest score whose key could exceed `key` (low bits cleared): key(v) > key  =>  v >= key_lower_score(key)
__device__ __forceinline__ float key_lower_score(uint32_t key, uint32_t low_mask) {
  uint32_t u = key & ~low_mask;
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  const float t = __uint_as_float(u);
  return (t == t) ? t : -INFINITY;  // a NaN bit pattern (keys of huge negative scores) must not reject everything
}

// sorted (descending) register list: insert x with a depth-2 network -- L'[i] = max(L[i], min(L[i-1], x)) -- instead of a
// 16-deep dependent min/max chain (x = 0 is a no-op: keys are >= 0)
template <int KP>
__device__ __forceinline__ void topk_insert(uint32_t (&L)[KP], uint32_t x) {
#pragma unroll
  for (int i = KP - 1; i >= 1; --i) L[i] = max(L[i], min(L[i - 1], x));
  L[0] = max(L[0], x);
}

// ---- max squared row norm of the DB (input of the certification bound) -------------------------------------
__global__ void __launch_bounds__(256) row_norm2_max_kernel(const float* __restrict__ db, int N, int D, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float best = 0.f;
  for (int r = warp; r < N; r += nwarps) {
    const float* p = db + (size_t)r * D;
    float ss = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float v = __ldg(p + c);
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    best = fmaxf(best, ss);
  }
  if (lane == 0 && best > 0.f) atomicMax(reinterpret_cast<int*>(out), __float_as_int(best));
}

// ---- scan ----------------------------------------------------------------------------------------------
#ifdef T2P_SCAN_TRACE  // tools/make_scan_trace.py: %globaltimer stamps of CTA (0, 0), read back with t2p_debug_scan_trace
__device__ unsigned long long scan_trace[128];
#define STR(i)                                                          \
  do {                                                                  \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (i) < 128) {              \
      unsigned long long t_;                                            \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));            \
      scan_trace[(i)] = t_;                                             \
    }                                                                   \
  } while (0)
#else
#define STR(i) do {} while (0)
#endif
// D[tmem] (+)= A[tmem] . B[smem]^T, kind::tf32 (A: one 32-bit column per K element)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  int spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    if (++spins > 400) __trap();
  }
}
__device__ __forceinline__ bool scan_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit_addr(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void scan_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
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
constexpr int TC_MAX_STAGES = 12;

struct ScanSmem {
  uint64_t q_full;
  uint64_t full[TC_MAX_STAGES];
  uint64_t empty[TC_MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_slot;
};

struct ScanParams {
  int B, N, D;
  int tile_n;      // DB rows per tile (UMMA N), multiple of 16
  int q_box_rows;  // rows of one query TMA box
  int db_box_rows; // rows of one DB TMA box (= min(tile_n, N))
  int q_pitch;     // bytes between the 32-float chunks of the query tile in shared memory (multiple of 1024)
  int st_pitch;    // bytes per pipeline stage (multiple of 1024)
  int stages;
  int nb_bits;     // width of the local-row field of the keys
  int dup;         // B <= 64: the query tile is loaded twice (TMEM lanes 64..127 mirror 0..63) and the two halves of
                   // the epilogue warps split the columns of every tile; they write separate key lists (2G sources)
  int qstream;     // B > 128 and one DB tile per CTA: the DB tile stays resident (q_pitch = its chunk pitch) and the QUERY
                   // tiles stream through the ring (st_pitch = 16 KB); one key list per (query, CTA) and query tile
  int qtiles;
  int a_tmem;      // not qstream: the query tile lives in TENSOR memory (A operand of the TS form; columns [256, 256 + D)), written
                   // once by the epilogue warps -- all of shared memory is the DB ring (the 128 KB tile left room for 2 stages)
  int slack;       // bytes after the ring the UMMA may read past a short resident tile (SS forms only)
  int epi_warps;   // 4 or 8 epilogue warps (1 or 2 key lists per (query, CTA, half))
  int debug;       // knock-outs of the trace build (T2P_SCAN_DEBUG; results are wrong): 1 = no UMMAs, 2 = no column scan;
                   // always 0 in the product library
};

template <int KP>
__global__ void __launch_bounds__(TC_THREADS, 1)
retrieve_scan_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_db,
                        const ScanParams p, const float* __restrict__ q_rows, uint32_t* __restrict__ part_keys) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nch = p.D >> 5;  // 128-byte chunks along K
  uint8_t* q_smem = base;
  uint8_t* st_smem = base + (size_t)nch * p.q_pitch;
  // the UMMA reads 128 rows (16 KB) from every query chunk whatever q_pitch is: 16 KB of slack follow the stages
  uint32_t* cand_smem = reinterpret_cast<uint32_t*>(st_smem + (size_t)p.stages * p.st_pitch + p.slack);
  ScanSmem* sm = reinterpret_cast<ScanSmem*>(st_smem + (size_t)p.stages * p.st_pitch + p.slack + TC_CAND_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, c = blockIdx.x;
  const int q0 = blockIdx.y * TC_QM;
  if (threadIdx.x == 0) STR(0);
  const int tile_n = p.tile_n;
  const int tiles = (p.N + tile_n - 1) / tile_n;
  // iterations of the pipeline: DB tiles of this CTA (query tile resident) or query tiles (qstream: DB tile resident)
  const int my_tiles = p.qstream ? p.qtiles : (tiles - c + G - 1) / G;  // host guarantees c < tiles

  if (threadIdx.x == 0) {
    mbar_init(&sm->q_full, p.a_tmem ? p.epi_warps : 1);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&sm->full[s], 1);
      mbar_init(&sm->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sm->tmem_full[a], 1);
      mbar_init(&sm->tmem_empty[a], p.epi_warps);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<TC_TMEM_COLS>(&sm->tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = sm->tmem_slot;
  if (threadIdx.x == 0) STR(1);

  auto WAIT = [&](uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); };  // parked waits: polling measured slower
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q);
      tma_prefetch_desc(&tmap_db);
      int stage = 0;
      uint32_t ph = 0;
      if (p.qstream) {
        // resident operand = this CTA's DB tile (all K chunks), streamed operand = the query tiles
        mbar_expect_tx(&sm->q_full, (uint32_t)(nch * p.db_box_rows * 128));
        for (int kc = 0; kc < nch; ++kc) tma_load_2d(q_smem + (size_t)kc * p.q_pitch, &tmap_db, kc * 32, c * tile_n, &sm->q_full);
        for (int qt = 0; qt < my_tiles; ++qt) {
          for (int kc = 0; kc < nch; ++kc) {
            WAIT(&sm->empty[stage], ph ^ 1);
            mbar_expect_tx(&sm->full[stage], (uint32_t)(p.q_box_rows * 128));
            tma_load_2d(st_smem + (size_t)stage * p.st_pitch, &tmap_q, kc * 32, qt * TC_QM, &sm->full[stage]);
            if (++stage == p.stages) { stage = 0; ph ^= 1; }
          }
        }
      } else {
        if (!p.a_tmem) {
          const int q_copies = p.dup ? 2 : 1;
          mbar_expect_tx(&sm->q_full, (uint32_t)(nch * q_copies * p.q_box_rows * 128));
          for (int kc = 0; kc < nch; ++kc) {
            tma_load_2d(q_smem + (size_t)kc * p.q_pitch, &tmap_q, kc * 32, q0, &sm->q_full);
            if (p.dup) tma_load_2d(q_smem + (size_t)kc * p.q_pitch + 64 * 128, &tmap_q, kc * 32, q0, &sm->q_full);
          }
        }
        for (int lt = 0; lt < my_tiles; ++lt) {
          const int row0 = (c + lt * G) * tile_n;
          for (int kc = 0; kc < nch; ++kc) {
            WAIT(&sm->empty[stage], ph ^ 1);
            mbar_expect_tx(&sm->full[stage], (uint32_t)(p.db_box_rows * 128));
            tma_load_2d(st_smem + (size_t)stage * p.st_pitch, &tmap_db, kc * 32, row0, &sm->full[stage]);
            if (++stage == p.stages) { stage = 0; ph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // One thread feeds the tensor pipe, and it shares its scheduler with an epilogue warp: every instruction between two
    // tcgen05.mma is on the critical path (the first version rebuilt descriptors and re-tested the mode per K chunk: ~900
    // cycles per chunk against 256 of MMA time).  The common mode (query tile in tensor memory) gets its own lean loop.
    // The whole warp runs the (warp-uniform) control flow and one ELECTED lane issues the tcgen05 instructions: inside a
    // per-thread branch (`if (lane == 0)`) the compiler wraps every UTCHMMA in an elect / branch loop.
    if (p.a_tmem && !(p.debug & 1)) {
      const uint32_t idesc = umma_idesc_tf32(TC_QM, tile_n);
      const uint64_t desc0 = umma_desc_sw128_kmajor(smem_u32(st_smem));
      const uint32_t pitch16 = (uint32_t)p.st_pitch >> 4;  // the address field of the descriptor counts 16-byte units
      const uint32_t a0 = tmem_base + TC_Q_COL;
      const uint32_t full0 = smem_u32(&sm->full[0]), empty0 = smem_u32(&sm->empty[0]);
      const int stages = p.stages;
      mbar_wait(&sm->q_full, 0);
      if (lane == 0) STR(3);
      int stage = 0;
      uint32_t ph = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int acc = lt & 1;
        mbar_wait(&sm->tmem_empty[acc], ((lt >> 1) & 1) ^ 1);
        const uint32_t d_tmem = tmem_base + acc * TC_TILE_MAX;
        for (int kc = 0; kc < nch; ++kc) {
          mbar_wait_addr(full0 + 8 * stage, ph);
          tc_fence_after_sync();
          if (scan_elect_one()) {
            const uint64_t b_desc = desc0 + (uint64_t)(stage * pitch16);
            const uint32_t a_col = a0 + kc * 32;
            umma_tf32_ts(d_tmem, a_col, b_desc, idesc, kc != 0);
            umma_tf32_ts(d_tmem, a_col + 8, b_desc + 2, idesc, true);
            umma_tf32_ts(d_tmem, a_col + 16, b_desc + 4, idesc, true);
            umma_tf32_ts(d_tmem, a_col + 24, b_desc + 6, idesc, true);
            umma_commit_addr(empty0 + 8 * stage);
            if (kc == nch - 1) umma_commit(&sm->tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; ph ^= 1; }
        }
        if (lane == 0) STR(8 + lt);
      }
    } else if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(TC_QM, tile_n);
      const uint32_t q_addr = smem_u32(q_smem), st_addr = smem_u32(st_smem);
      WAIT(&sm->q_full, 0);
      STR(3);
      int stage = 0;
      uint32_t ph = 0;
      {
        for (int lt = 0; lt < my_tiles; ++lt) {
          const int acc = lt & 1;
          WAIT(&sm->tmem_empty[acc], ((lt >> 1) & 1) ^ 1);
          tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + acc * TC_TILE_MAX;
          for (int kc = 0; kc < nch; ++kc) {
            WAIT(&sm->full[stage], ph);
            tc_fence_after_sync();
            // A = queries (M = 128 TMEM lanes), B = DB rows (N = tile_n columns); which of them is resident depends on the mode
            const uint64_t res_desc = umma_desc_sw128_kmajor(q_addr + kc * p.q_pitch);
            const uint64_t str_desc = umma_desc_sw128_kmajor(st_addr + stage * p.st_pitch);
            const uint64_t a_desc = p.qstream ? str_desc : res_desc;
            const uint64_t b_desc = p.qstream ? res_desc : str_desc;
            if (p.debug & 1) {
            } else if (p.a_tmem) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)  // A: 8 tensor-memory columns per instruction
                umma_tf32_ts(d_tmem, tmem_base + TC_Q_COL + kc * 32 + ks * 8, str_desc + 2 * ks, idesc, (kc | ks) != 0);
            } else {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)  // K = 8 tf32 = 32 bytes per instruction: +2 in the 16-byte address field
                umma_tf32_ss(d_tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc, (kc | ks) != 0);
            }
            umma_commit(&sm->empty[stage]);
            if (++stage == p.stages) { stage = 0; ph ^= 1; }
          }
          umma_commit(&sm->tmem_full[acc]);
          STR(8 + lt);
        }
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: TMEM lane = query, columns = DB rows of the tile =====
    // A lane keeps the KP best keys of its query as a sorted register list.  Scores are first filtered against the lane's
    // threshold (a float lower bound of its KP-th key; one compare per score) and the survivors are appended, as keys, to the
    // lane's column of a shared-memory buffer; when some lane's buffer runs full the warp merges the buffers into the lists
    // (max-over-lanes rounds of the depth-2 insertion network) and refreshes the thresholds.  The warp-wide insertion per
    // column of the first version ran for ~90 % of the columns (32 queries share the branch) and was a 16-deep dependent chain.
    const int quad = warp & 3;  // TMEM lanes [32*quad, 32*quad+32) are the ones this warp may read
    const int tl = quad * 32 + lane;
    const int half = p.dup ? (tl >> 6) : 0;
    const int sub = (warp - 2) >> 2;  // which of the warps of the quadrant
    const int subs = p.epi_warps >> 2;
    const uint32_t low_mask = (1u << p.nb_bits) - 1u;
    if (p.a_tmem) {
      // The query tile goes to tensor memory once: TMEM lane = query row (rows 64.. mirror 0..63 in dup mode), column = channel.
      // A warp owns the 32 rows of its lane quadrant and moves them in blocks of 32 channels: coalesced 16-byte loads (8 lanes
      // per row), a transpose through the warp's 4 KB of shared memory (16-byte units XOR-swizzled by the row: conflict-free
      // both ways), then lane = row reads its 32 channels back for one tcgen05.st.  (A lane reading its own row straight from
      // global memory costs 32 L1 wavefronts per load instruction: 5 us per CTA.)  The warps of a quadrant take alternate blocks.
      const uint32_t xbuf = smem_u32(cand_smem + (warp - 2) * (TC_CAND_CAP * 32));
      const int row_base = q0 + (p.dup ? (quad & 1) * 32 : quad * 32);
      const int unit = lane & 7, rsub = lane >> 3;
      auto load_block = [&](float4 (&x)[8], int c0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int qr = row_base + 4 * i + rsub;
          x[i] = qr < p.B ? __ldg(reinterpret_cast<const float4*>(q_rows + (size_t)qr * p.D + c0) + unit) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      const int step = 32 * subs;
      float4 x[8];
      int c0 = 32 * sub;
      if (c0 < p.D) load_block(x, c0);
      for (; c0 < p.D; c0 += step) {
        float4 nx[8];
        const bool more = c0 + step < p.D;
        if (more) load_block(nx, c0 + step);  // in flight while this block is transposed
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 4 * i + rsub;
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(xbuf + r * 128 + ((unit ^ (r & 7)) << 4)), "f"(x[i].x),
                       "f"(x[i].y), "f"(x[i].z), "f"(x[i].w));
        }
        __syncwarp();
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v[4 * j]), "=r"(v[4 * j + 1]), "=r"(v[4 * j + 2]), "=r"(v[4 * j + 3])
                       : "r"(xbuf + lane * 128 + ((j ^ (lane & 7)) << 4)));
        scan_tmem_st32(tmem_base + ((uint32_t)(quad * 32) << 16) + TC_Q_COL + c0, v);
        if (more) {
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = nx[i];
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->q_full);
      if (warp == 2 && lane == 0) STR(2);
    }
    // the lane's column of the warp's candidate buffer, as a shared-space address (slot r at +128 r bytes)
    const uint32_t buf = smem_u32(cand_smem + (warp - 2) * (TC_CAND_CAP * 32) + lane);
    uint32_t L[KP];
#pragma unroll
    for (int i = 0; i < KP; ++i) L[i] = 0u;
    float thr = -INFINITY;
    uint32_t cnt = 0;
    auto merge = [&]() {
      const uint32_t rounds = __reduce_max_sync(0xffffffffu, cnt);
      for (uint32_t r = 0; r < rounds; ++r) {
        uint32_t x;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(buf + (r << 7)));
        topk_insert<KP>(L, r < cnt ? x : 0u);
      }
      cnt = 0;
      if (L[KP - 1] != 0u) thr = key_lower_score(L[KP - 1], low_mask);
    };
    const int n16 = tile_n >> 4;
    const int ch_begin = (p.dup && half) ? ((n16 + 1) >> 1) : 0;
    const int ch_end = (p.dup && !half) ? ((n16 + 1) >> 1) : n16;
    const int parts = (p.dup ? 2 : 1) * subs;  // key lists per (query, CTA)
    const int nsrc = parts * G;
    const int src = parts * c + half * subs + sub;
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int acc = lt & 1;
      // qstream: iteration lt is query tile lt against the CTA's single DB tile; otherwise DB tile lt against query tile blockIdx.y
      const int qbase = p.qstream ? lt * TC_QM : q0;
      const int qrow = qbase + (p.dup ? (tl & 63) : tl);
      const bool warp_active = (qbase + (p.dup ? ((quad & 1) * 32) : quad * 32)) < p.B;
      const int row0 = p.qstream ? c * tile_n : (c + lt * G) * tile_n;
      const uint32_t local0 = p.qstream ? 0u : (uint32_t)(lt * TC_TILE_MAX);
      const int ncols = min(tile_n, p.N - row0);  // valid columns of this tile (the rows TMA zero-fills must not become candidates)
      mbar_wait(&sm->tmem_full[acc], (lt >> 1) & 1);
      if (warp == 2 && lane == 0) STR(40 + lt);
      tc_fence_after_sync();
      if (warp_active && ch_begin + sub < ch_end && !(p.debug & 2)) {
        const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TC_TILE_MAX;
        // branch-free filter: the key of every score is written to the lane's next free slot, the slot is only kept (cnt
        // advances) if the score passes the threshold.  Four columns per group: their slots are cnt + (prefix of the hit
        // bits), so the dependent chain through cnt is one add per four columns.
        auto scan16_impl = [&](const uint32_t (&v)[16], int ch, auto full_tag) {
          constexpr bool FULL = decltype(full_tag)::value;
          const uint32_t key0 = low_mask - (local0 + (uint32_t)(ch * 16));
          const int left = ncols - ch * 16;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t key[4], hit[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = 4 * g + i;
              uint32_t u = v[j];
              u ^= (uint32_t)((int32_t)u >> 31) | 0x80000000u;  // order-preserving bits (score_to_key)
              key[i] = (u & ~low_mask) | (key0 - (uint32_t)j);
              hit[i] = (__uint_as_float(v[j]) >= thr && (FULL || j < left)) ? 1u : 0u;
            }
            const uint32_t s1 = hit[0], s2 = hit[0] + hit[1], s3 = s2 + hit[2], s4 = s3 + hit[3];
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(buf + (cnt << 7)), "r"(key[0]));
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(buf + ((cnt + s1) << 7)), "r"(key[1]));
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(buf + ((cnt + s2) << 7)), "r"(key[2]));
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(buf + ((cnt + s3) << 7)), "r"(key[3]));
            cnt += s4;
          }
          if (__any_sync(0xffffffffu, cnt > (uint32_t)(TC_CAND_CAP - 16))) merge();
        };
        const bool full_tile = ncols == tile_n;  // every tile but the last one of the DB
        auto scan16 = [&](const uint32_t (&v)[16], int ch) {
          if (full_tile) scan16_impl(v, ch, std::true_type{});
          else scan16_impl(v, ch, std::false_type{});
        };
        // the TMEM load of the warp's next chunk is in flight while the current one is filtered
        uint32_t va[16], vb[16];
        int ch = ch_begin + sub;
        tmem_ld_32x16(t_addr + ch * 16, va);
        tmem_ld_wait_on(va);
        while (true) {
          const bool more = ch + subs < ch_end;
          if (more) tmem_ld_32x16(t_addr + (ch + subs) * 16, vb);
          scan16(va, ch);
          if (!more) break;
          tmem_ld_wait_on(vb);
          const bool more2 = ch + 2 * subs < ch_end;
          if (more2) tmem_ld_32x16(t_addr + (ch + 2 * subs) * 16, va);
          scan16(vb, ch + subs);
          if (!more2) break;
          tmem_ld_wait_on(va);
          ch += 2 * subs;
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->tmem_empty[acc]);
      if (warp == 2 && lane == 0) STR(72 + lt);
      if (p.qstream || lt == my_tiles - 1) {  // the list of (query, CTA) is complete: write it (qstream: one per query tile)
        if (warp_active) merge();
        if (qrow < p.B && warp_active) {
          uint4* dst = reinterpret_cast<uint4*>(part_keys + ((size_t)qrow * nsrc + src) * KP);
#pragma unroll
          for (int i = 0; i < KP / 4; ++i) dst[i] = make_uint4(L[4 * i], L[4 * i + 1], L[4 * i + 2], L[4 * i + 3]);
        }
#pragma unroll
        for (int i = 0; i < KP; ++i) L[i] = 0u;
        thr = -INFINITY;
      }
    }
  }
  if (warp == 2 && lane == 0) STR(4);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 1) tmem_dealloc<TC_TMEM_COLS>(tmem_base);
  if (threadIdx.x == 0) STR(5);
}

// ---- select ----------------------------------------------------------------------------------------------
// float64 dot product of one query and one DB row by one warp (lanes stride the channels, fixed xor tree)
__device__ __forceinline__ double warp_dot_f64(const float* __restrict__ q_smem, const float* __restrict__ d, int D, int lane) {
  double acc = 0.0;
  for (int ch = lane; ch < D; ch += 32) acc = fma((double)q_smem[ch], (double)__ldg(d + ch), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

constexpr int SEL_SURV_MAX = 256;

struct __align__(16) SelSmem {
  unsigned wsum[2][SEL_THREADS / 32];
  unsigned n_cand, n_surv, below_max, m_last, fallback, tau0;
  double tk, qq;
  uint32_t surv_key[SEL_SURV_MAX];
  int32_t surv_pos[SEL_SURV_MAX];
  uint32_t cand_key[SEL_CAND_MAX];
  int32_t cand_row[SEL_CAND_MAX];
  double cand_score[SEL_CAND_MAX];
  double wl_s[8][32];
  int64_t wl_i[8][32];
};

struct SelParams {
  int B, N, D;
  int nsrc;    // key lists per query (CTAs of the scan x parts)
  int G;       // CTAs of the scan (row decoding)
  int parts;   // key lists per (query, CTA): 2 epilogue warps per lane quadrant, x2 in dup mode
  int KP, NC, k, tile_n, nb_bits;
  int force_rescan;
  int64_t idx_base;
};

// one query by the whole CTA (every thread of the CTA calls it with the same qi)
__device__ void select_one_query(const int qi, uint8_t* sel_raw, const float* __restrict__ q, const float* __restrict__ db,
                                 const SelParams& p, const uint32_t* __restrict__ part_keys,
                                 const float* __restrict__ db_norm2_max, double* __restrict__ out_s, int64_t* __restrict__ out_i,
                                 int32_t* __restrict__ stats) {
  SelSmem* sm = reinterpret_cast<SelSmem*>(sel_raw);
  float* qs = reinterpret_cast<float*>(sel_raw + sizeof(SelSmem));  // [D]
  uint32_t* keys = reinterpret_cast<uint32_t*>(qs + p.D);             // [nsrc*KP]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = p.D, N = p.N, KP = p.KP, NC = p.NC, k = p.k, nsrc = p.nsrc;
  const int total = nsrc * KP;
  const uint32_t low_mask = (1u << p.nb_bits) - 1u;
  const double NINF = __longlong_as_double(0xfff0000000000000LL);
  auto row_of = [&](int pos, uint32_t key) -> int {
    const int src = pos / KP;
    const int cta = src / p.parts;
    const uint32_t local = low_mask - (key & low_mask);
    return (cta + (int)(local >> 7) * p.G) * p.tile_n + (int)(local & 127u);
  };

  {
    const uint4* src = reinterpret_cast<const uint4*>(part_keys + (size_t)qi * total);  // KP % 4 == 0, 16-byte aligned
    uint4* dst = reinterpret_cast<uint4*>(keys);
    for (int t = tid; t < total / 4; t += SEL_THREADS) dst[t] = __ldg(src + t);
  }
  for (int t = tid; t < D; t += SEL_THREADS) qs[t] = __ldg(q + (size_t)qi * D + t);
  if (tid == 0) {
    sm->n_cand = 0;
    sm->n_surv = 0;
    sm->below_max = 0;
    sm->m_last = 0;
    sm->fallback = 0;
    sm->tau0 = 0;
    sm->tk = NINF;
    sm->qq = 0.0;
  }
  __syncthreads();

  // a full list may hide rows with keys up to its last entry
  {
    uint32_t ml = 0u;
    for (int t = tid; t < nsrc; t += SEL_THREADS) ml = max(ml, keys[t * KP + KP - 1]);
    ml = __reduce_max_sync(0xffffffffu, ml);
    if (lane == 0 && ml) atomicMax(&sm->m_last, ml);
  }

  // ---- pre-filter: a threshold tau0 with at least NC keys >= tau0 and typically only a few more.  List heads are dealt
  // round-robin to the 8 warps; each warp finds its NC/8-th largest head with a few redux rounds (removing all copies of
  // the current maximum per round keeps ">= NC/8 heads of this warp are >= the result" true under ties), tau0 = the
  // minimum over the warps.  O(NC/8) per warp instead of ranking every head against every other.
  constexpr int NW = SEL_THREADS / 32;
  const bool fast = nsrc <= 2 * SEL_THREADS && nsrc >= NC && (NC % NW) == 0;
  if (fast) {
    const int t0 = warp + NW * lane, t1 = t0 + SEL_THREADS;
    uint32_t v0 = t0 < nsrc ? keys[t0 * KP] : 0u;
    uint32_t v1 = t1 < nsrc ? keys[t1 * KP] : 0u;
    uint32_t m = 0u;
    for (int r = 0; r < NC / NW; ++r) {
      m = __reduce_max_sync(0xffffffffu, max(v0, v1));
      v0 = v0 == m ? 0u : v0;
      v1 = v1 == m ? 0u : v1;
    }
    if (lane == 0) sm->wsum[0][warp] = m;
    __syncthreads();
    if (tid == 0) {
      uint32_t t = 0xffffffffu;
#pragma unroll
      for (int w = 0; w < NW; ++w) t = min(t, sm->wsum[0][w]);
      sm->tau0 = t;  // 0 if some warp saw fewer than NC/8 non-empty heads: everything survives, the general path decides
    }
    __syncthreads();
  }
  const uint32_t tau0 = sm->tau0;  // 0: keep every non-empty key
  {
    uint32_t below = 0u;
    for (int t0 = 0; t0 < total; t0 += SEL_THREADS) {
      const int t = t0 + tid;
      const uint32_t x = t < total ? keys[t] : 0u;
      const bool keep = x != 0u && x >= tau0;
      if (!keep) below = max(below, x);
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (m) {
        unsigned basev = 0;
        const int leader = __ffs(m) - 1;
        if (lane == leader) basev = atomicAdd(&sm->n_surv, (unsigned)__popc(m));
        basev = __shfl_sync(0xffffffffu, basev, leader);
        const unsigned slot = basev + __popc(m & ((1u << lane) - 1u));
        if (keep && slot < SEL_SURV_MAX) {
          sm->surv_key[slot] = x;
          sm->surv_pos[slot] = t;
        }
      }
    }
    below = __reduce_max_sync(0xffffffffu, below);
    if (lane == 0 && below) atomicMax(&sm->below_max, below);
  }
  __syncthreads();
  const int n_surv = (int)sm->n_surv;
  int n_cand = 0;
  bool need_rescan = p.force_rescan != 0;

  if (n_surv <= NC + 8) {
    // few survivors: every one becomes a candidate (at least NC of them are the NC best keys overall)
    if (tid < n_surv) {
      const uint32_t mine = sm->surv_key[tid];
      sm->cand_key[tid] = mine;
      sm->cand_row[tid] = row_of(sm->surv_pos[tid], mine);
    }
    n_cand = n_surv;
  } else if (n_surv <= SEL_SURV_MAX) {
    // exact selection of the NC best survivors by rank counting: rank = #survivors that sort before (key desc, position asc)
    uint32_t drop = 0u;
    if (tid < n_surv) {
      const uint32_t mine = sm->surv_key[tid];
      const int mpos = sm->surv_pos[tid];
      int r = 0;
      for (int j = 0; j < n_surv; ++j) {
        const uint32_t h = sm->surv_key[j];
        r += (h > mine || (h == mine && sm->surv_pos[j] < mpos)) ? 1 : 0;
      }
      if (r < NC) {
        sm->cand_key[r] = mine;
        sm->cand_row[r] = row_of(mpos, mine);
      } else {
        drop = mine;
      }
    }
    drop = __reduce_max_sync(0xffffffffu, drop);
    if (lane == 0 && drop) atomicMax(&sm->below_max, drop);
    n_cand = NC;
  } else {
    // general path (many sources or dense keys): tau = NC-th largest key by bisection on the key bits
    uint32_t tau = 0u;
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = tau | (1u << bit);
      int cnt = 0;
      for (int t = tid; t < total; t += SEL_THREADS) cnt += (keys[t] >= cand) ? 1 : 0;
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      unsigned* ws = sm->wsum[bit & 1];
      if (lane == 0) ws[warp] = (unsigned)cnt;
      __syncthreads();
      unsigned tot = 0;
#pragma unroll
      for (int w = 0; w < SEL_THREADS / 32; ++w) tot += ws[w];
      if (tot >= (unsigned)NC) tau = cand;
    }
    uint32_t below = 0u;
    for (int t = tid; t < total; t += SEL_THREADS) {
      const uint32_t x = keys[t];
      if (x != 0u && x >= tau) {
        const unsigned slot = atomicAdd(&sm->n_cand, 1u);
        if (slot < SEL_CAND_MAX) {
          sm->cand_key[slot] = x;
          sm->cand_row[slot] = row_of(t, x);
        }
      } else {
        below = max(below, x);
      }
    }
    below = __reduce_max_sync(0xffffffffu, below);
    if (lane == 0 && below) atomicMax(&sm->below_max, below);
    __syncthreads();
    n_cand = (int)sm->n_cand;
    if (n_cand > SEL_CAND_MAX) need_rescan = true;  // more ties at tau than the candidate buffer holds
  }
  __syncthreads();

  if (!need_rescan) {
    // float64 re-scoring: each warp takes up to 8 candidates (f = warp + 8 j).  D % 128 == 0 (the tensor path needs
    // D % 32 == 0; 128, 256 are the model sizes): a row is read as float4, 32 lanes x D/128 loads, and ALL loads of all the
    // warp's candidates are issued before the first use -- one L2/DRAM round trip instead of one per 32 channels.
    {
      constexpr int NWS = SEL_THREADS / 32, PER = SEL_CAND_MAX / NWS;
      if ((D & 127) == 0 && D <= 256) {
        const int nv = D >> 7;  // float4 loads per lane and row: 1 or 2
        const float4* q4 = reinterpret_cast<const float4*>(qs);
        float4 qv[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) qv[c] = c < nv ? q4[lane + 32 * c] : make_float4(0.f, 0.f, 0.f, 0.f);
        constexpr int HALF = PER / 2;  // 4 candidates per warp and pass: <= 32 candidates (the usual case) need one pass
        for (int pass = 0; pass < 2; ++pass) {
          if (warp + NWS * (pass * HALF) >= n_cand) break;  // warp-uniform
          float4 v[HALF][2];
#pragma unroll
          for (int jx = 0; jx < HALF; ++jx) {
            const int f = warp + NWS * (pass * HALF + jx);
            const float4* row = reinterpret_cast<const float4*>(db + (size_t)sm->cand_row[f < n_cand ? f : 0] * D);
#pragma unroll
            for (int c = 0; c < 2; ++c)
              v[jx][c] = (f < n_cand && c < nv) ? __ldg(row + lane + 32 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int jx = 0; jx < HALF; ++jx) {
            const int f = warp + NWS * (pass * HALF + jx);
            if (f < n_cand) {  // warp-uniform
              double a = 0.0;
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                a = fma((double)qv[c].x, (double)v[jx][c].x, a);
                a = fma((double)qv[c].y, (double)v[jx][c].y, a);
                a = fma((double)qv[c].z, (double)v[jx][c].z, a);
                a = fma((double)qv[c].w, (double)v[jx][c].w, a);
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
              if (lane == 0) sm->cand_score[f] = a;
            }
          }
        }
      } else {
        double acc[PER];
        const float* rows[PER];
#pragma unroll
        for (int jx = 0; jx < PER; ++jx) {
          const int f = warp + NWS * jx;
          acc[jx] = 0.0;
          rows[jx] = db + (size_t)sm->cand_row[f < n_cand ? f : 0] * D;
        }
        for (int ch = lane; ch < D; ch += 32) {
          const double qv = (double)qs[ch];
#pragma unroll
          for (int jx = 0; jx < PER; ++jx)
            if (warp + NWS * jx < n_cand) acc[jx] = fma(qv, (double)__ldg(rows[jx] + ch), acc[jx]);
        }
#pragma unroll
        for (int jx = 0; jx < PER; ++jx) {
          const int f = warp + NWS * jx;
          if (f < n_cand) {  // warp-uniform
            double a = acc[jx];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) sm->cand_score[f] = a;
          }
        }
      }
    }
    __syncthreads();
    // rank by (score desc, row asc); the thread holding rank k-1 publishes the k-th best exact score
    int my_rank = -1;
    if (tid < n_cand) {
      const double ms = sm->cand_score[tid];
      const int mr = sm->cand_row[tid];
      int r = 0;
      for (int g = 0; g < n_cand; ++g) {
        const double gs = sm->cand_score[g];
        const int gr = sm->cand_row[g];
        r += (gs > ms || (gs == ms && gr < mr)) ? 1 : 0;
      }
      my_rank = r;
      if (r == k - 1) sm->tk = ms;
    }
    if (warp == SEL_THREADS / 32 - 1) {  // |q|^2 for the error bound
      double qq = 0.0;
      for (int ch = lane; ch < D; ch += 32) qq = fma((double)qs[ch], (double)qs[ch], qq);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
      if (lane == 0) sm->qq = qq;
    }
    __syncthreads();
    // certification: every non-candidate row has a TF32 score <= key_upper_score(ukey), hence an exact score
    // <= that + eps*|q|*max|d|; the result stands only if the k-th exact score is strictly above this bound
    if (tid == 0) {
      const uint32_t ukey = max(sm->below_max, sm->m_last);
      bool certified = true;
      if (ukey != 0u) {
        if (n_cand < k) {
          certified = false;
        } else {
          const double bound = (double)key_upper_score(ukey, low_mask) +
                               TC_EPS * sqrt(sm->qq) * sqrt((double)__ldg(db_norm2_max) * (1.0 + 1e-5));
          certified = sm->tk > bound;
        }
      }
      sm->fallback = certified ? 0u : 1u;
    }
    __syncthreads();
    need_rescan = sm->fallback != 0u;
    if (!need_rescan) {
      if (tid < n_cand && my_rank < k) {
        out_s[(size_t)qi * k + my_rank] = sm->cand_score[tid];
        out_i[(size_t)qi * k + my_rank] = p.idx_base + (int64_t)sm->cand_row[tid];
      }
      for (int r = n_cand + tid; r < k; r += SEL_THREADS) {  // N < k: pad
        out_s[(size_t)qi * k + r] = NINF;
        out_i[(size_t)qi * k + r] = -1;
      }
      if (tid == 0 && stats) atomicAdd(stats + 0, 1);
      return;
    }
  }

  // exact float64 rescan of the whole DB for this query (rare: dense ties / clustered scores / forced)
  {
    const int64_t IMAX = 0x7fffffffffffffffLL;
    WarpTopK<double, int64_t> top;
    top.init(k, NINF, IMAX);
    for (int r0 = warp * 32; r0 < N; r0 += SEL_THREADS) {  // each warp: 32 consecutive rows per round, lane l keeps row r0+l
      double mine = NINF;
      for (int rr = 0; rr < 32; ++rr) {
        const int r = r0 + rr;
        if (r < N) {  // warp-uniform
          const double s = warp_dot_f64(qs, db + (size_t)r * D, D, lane);
          if (lane == rr) mine = s;
        }
      }
      top.offer(r0 + lane < N, mine, (int64_t)(r0 + lane));
    }
    sm->wl_s[warp][lane] = top.s;
    sm->wl_i[warp][lane] = top.i;
    __syncthreads();
    if (warp == 0) {
      for (int w = 1; w < SEL_THREADS / 32; ++w) top.offer(lane < k && sm->wl_i[w][lane] != IMAX, sm->wl_s[w][lane], sm->wl_i[w][lane]);
      if (lane < k) {
        const bool ok = top.i != IMAX;
        out_s[(size_t)qi * k + lane] = ok ? top.s : NINF;
        out_i[(size_t)qi * k + lane] = ok ? p.idx_base + top.i : (int64_t)-1;
      }
      if (lane == 0 && stats) atomicAdd(stats + 1, 1);
    }
  }
}

// grid = B CTAs, one query each (only_flagged == nullptr), or -- second stage behind retrieve_select_warp_kernel -- a small grid
// whose CTAs walk over the queries and redo only the ones the warp kernel could not certify (usually none: the launch then costs
// a handful of CTAs reading their flags instead of B CTAs spread over every SM)
__global__ void __launch_bounds__(SEL_THREADS, 3)
retrieve_select_kernel(const float* __restrict__ q, const float* __restrict__ db, const SelParams p,
                       const uint32_t* __restrict__ part_keys, const float* __restrict__ db_norm2_max,
                       double* __restrict__ out_s, int64_t* __restrict__ out_i, int32_t* __restrict__ stats,
                       const int32_t* __restrict__ only_flagged) {
  extern __shared__ __align__(16) uint8_t sel_raw[];
  for (int qi = blockIdx.x; qi < p.B; qi += gridDim.x) {
    if (only_flagged != nullptr && only_flagged[qi] == 0) continue;  // CTA-uniform
    select_one_query(qi, sel_raw, q, db, p, part_keys, db_norm2_max, out_s, out_i, stats);
    __syncthreads();  // shared memory is reused by the next query
  }
}

// ---- select, one WARP per query ------------------------------------------------------------------------------------------
// The per-query work of the select is tiny (a few hundred keys, <= 32 rows to re-score); a 256-thread CTA per query spends
// its time in barriers and occupies an SM slot for ~17 us.  Here a warp does a whole query: the sorted key lists of the scan
// CTAs go to the warp's slice of shared memory, lane l owns lists l, l + 32, ..., and the best keys are popped one per round
// (redux over the lanes' best heads); the first 16 candidates are re-scored in float64 (all row loads of 8 candidates in
// flight), ranked with shuffles and CERTIFIED against the best remaining key exactly like retrieve_select_kernel; if that
// fails the next 16 are added; a query that still cannot be certified is flagged for the CTA-per-query kernel (which also
// owns the exact rescan).
constexpr int SELW_WARPS = 2;   // queries per CTA: the kernel is instruction-bound per warp (~6k instructions per query), so the
                                // warps are spread over many SMs (16 per CTA: 26 us for 64 queries on 4 SMs)
constexpr int SELW_MAX_LPL = 10;  // lists per lane: nsrc <= 320

__global__ void __launch_bounds__(32 * SELW_WARPS)
retrieve_select_warp_kernel(const float* __restrict__ q, const float* __restrict__ db, const SelParams p,
                            const uint32_t* __restrict__ part_keys, const float* __restrict__ db_norm2_max,
                            double* __restrict__ out_s, int64_t* __restrict__ out_i, int32_t* __restrict__ stats,
                            int32_t* __restrict__ flagged) {
  extern __shared__ __align__(16) uint8_t selw_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.x * SELW_WARPS + warp;
  if (qi >= p.B) return;
  const int D = p.D, KP = p.KP, k = p.k, nsrc = p.nsrc;
  const int total = nsrc * KP;
  uint32_t* keys = reinterpret_cast<uint32_t*>(selw_raw) + (size_t)warp * total;  // [nsrc][KP], each list sorted descending
  const uint32_t low_mask = (1u << p.nb_bits) - 1u;
  const double NINF = __longlong_as_double(0xfff0000000000000LL);
  {
    const uint4* src = reinterpret_cast<const uint4*>(part_keys + (size_t)qi * total);
    uint4* dst = reinterpret_cast<uint4*>(keys);
    for (int t = lane; t < total / 4; t += 32) dst[t] = __ldg(src + t);
  }
  // the query in registers (D % 128 == 0, D <= 256 on this path) and |q|^2
  const int nv = D >> 7;
  float4 qv[2];
  double qq = 0.0;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    qv[c] = c < nv ? __ldg(reinterpret_cast<const float4*>(q + (size_t)qi * D) + lane + 32 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    qq = fma((double)qv[c].x, (double)qv[c].x, qq);
    qq = fma((double)qv[c].y, (double)qv[c].y, qq);
    qq = fma((double)qv[c].z, (double)qv[c].z, qq);
    qq = fma((double)qv[c].w, (double)qv[c].w, qq);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
  __syncwarp();

  // a full list may hide rows with keys up to its last entry
  uint32_t m_last = 0u;
  for (int l = lane; l < nsrc; l += 32) m_last = max(m_last, keys[l * KP + KP - 1]);
  m_last = __reduce_max_sync(0xffffffffu, m_last);

  // cursors of this lane's lists, 5 bits each (0..KP <= 32 needs 6 bits for KP = 32: use 6)
  unsigned long long cur = 0ull;
  auto best_of_lane = [&](uint32_t& bkey, int& bpos) {
    bkey = 0u;
    bpos = 0;
#pragma unroll
    for (int i = 0; i < SELW_MAX_LPL; ++i) {
      if (32 * i >= nsrc) break;  // warp-uniform: no lane owns a list beyond this
      const int l = lane + 32 * i;
      const int c = (int)((cur >> (6 * i)) & 63ull);
      if (l < nsrc && c < KP) {
        const uint32_t x = keys[l * KP + c];
        if (x > bkey) { bkey = x; bpos = l * KP + c; }
      }
    }
  };
  uint32_t bkey;
  int bpos;
  best_of_lane(bkey, bpos);

  auto row_of = [&](int pos, uint32_t key) -> int {
    const int src = pos / KP;
    const int cta = src / p.parts;
    const uint32_t local = low_mask - (key & low_mask);
    return (cta + (int)(local >> 7) * p.G) * p.tile_n + (int)(local & 127u);
  };

  // candidate r lives in lane r
  int my_row = 0;
  double my_score = NINF;
  int n_cand = 0;
  bool certified = false;
  double tk = NINF;
  int my_rank = 64;
  for (int round = 0; round < 2 && !certified; ++round) {
    // ---- pop the next 16 best keys ----
    const int first = n_cand;
    for (int r = first; r < first + 16; ++r) {
      const uint32_t m = __reduce_max_sync(0xffffffffu, bkey);
      if (m == 0u) break;  // warp-uniform: nothing left
      const unsigned who = __ballot_sync(0xffffffffu, bkey == m);
      const int winner = __ffs(who) - 1;
      const int wpos = __shfl_sync(0xffffffffu, bpos, winner);
      if (lane == r) my_row = row_of(wpos, m);
      if (lane == winner) {
        const int i = (wpos / KP - lane) >> 5;
        cur += 1ull << (6 * i);
        best_of_lane(bkey, bpos);
      }
      n_cand = r + 1;
    }
    if (n_cand == first) break;
    // ---- float64 re-scoring of candidates [first, n_cand): 8 at a time, all row loads in flight ----
    for (int g0 = first; g0 < n_cand; g0 += 8) {
      float4 v[8][2];
#pragma unroll
      for (int jx = 0; jx < 8; ++jx) {
        const int f = g0 + jx;
        const int row = __shfl_sync(0xffffffffu, my_row, f & 31);
        const float4* rp = reinterpret_cast<const float4*>(db + (size_t)(f < n_cand ? row : 0) * D);
#pragma unroll
        for (int c = 0; c < 2; ++c) v[jx][c] = (f < n_cand && c < nv) ? __ldg(rp + lane + 32 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int jx = 0; jx < 8; ++jx) {
        const int f = g0 + jx;
        double a = 0.0;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          a = fma((double)qv[c].x, (double)v[jx][c].x, a);
          a = fma((double)qv[c].y, (double)v[jx][c].y, a);
          a = fma((double)qv[c].z, (double)v[jx][c].z, a);
          a = fma((double)qv[c].w, (double)v[jx][c].w, a);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (f < n_cand && lane == f) my_score = a;
      }
    }
    // ---- rank by (score desc, row asc) ----
    int r = 0;
    for (int g = 0; g < n_cand; ++g) {
      const double gs = __shfl_sync(0xffffffffu, my_score, g);
      const int gr = __shfl_sync(0xffffffffu, my_row, g);
      r += (gs > my_score || (gs == my_score && gr < my_row)) ? 1 : 0;
    }
    my_rank = lane < n_cand ? r : 64;
    const unsigned kth = __ballot_sync(0xffffffffu, my_rank == k - 1);
    tk = kth ? __shfl_sync(0xffffffffu, my_score, __ffs(kth) - 1) : NINF;
    // ---- certification (same bound as retrieve_select_kernel) ----
    const uint32_t ukey = max(__reduce_max_sync(0xffffffffu, bkey), m_last);
    if (ukey == 0u) {
      certified = true;  // every row of the DB is a candidate
    } else if (n_cand >= k) {
      const double bound = (double)key_upper_score(ukey, low_mask) +
                           TC_EPS * sqrt(qq) * sqrt((double)__ldg(db_norm2_max) * (1.0 + 1e-5));
      certified = tk > bound;
    }
  }
  if (certified) {
    if (my_rank < k) {
      out_s[(size_t)qi * k + my_rank] = my_score;
      out_i[(size_t)qi * k + my_rank] = p.idx_base + (int64_t)my_row;
    }
    for (int r = n_cand + lane; r < k; r += 32) {  // N < k: pad
      out_s[(size_t)qi * k + r] = NINF;
      out_i[(size_t)qi * k + r] = -1;
    }
    if (lane == 0) {
      flagged[qi] = 0;
      if (stats) atomicAdd(stats + 0, 1);
    }
  } else if (lane == 0) {
    flagged[qi] = 1;  // the CTA-per-query kernel redoes this query (wider candidate set, exact rescan if needed)
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// [rows, D] float32 row-major, box = [box_rows x 32 floats], 128-byte swizzle, out-of-bounds rows read as zero
static int make_tmap_rows(CUtensorMap* m, const float* ptr, int rows, int D, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  T2P_REQUIRE(fn != nullptr, T2P_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)D * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  T2P_REQUIRE(r == CUDA_SUCCESS, T2P_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%d D=%d box_rows=%d)", (int)r,
              rows, D, box_rows);
  return T2P_OK;
}

static int ceil_log2(int x) {
  int b = 0;
  while ((1 << b) < x) ++b;
  return b;
}

// epilogue warps of the scan: 8 halve the latency of a scan, 4 cost fewer threads and half the key lists for the select
// (T2P_SCAN_EPI overrides, for experiments)
static int scan_epi_warps() {
  static const int n = [] { const char* e = getenv("T2P_SCAN_EPI"); return (e && atoi(e) == 4) ? 4 : (e && atoi(e) == 8) ? 8 : TC_EPI_DEFAULT; }();
  return n;
}

TcPlan tc_plan(int B, int N, int D, int k, int sms) {
  TcPlan p = {};
  p.ok = false;
  if (D % 32 != 0 || D < 32 || D > 256 || k < 1 || k > 26 || N < 1 || B < 1) return p;
  p.KP = k <= 16 ? 16 : 32;
  p.NC = k <= 16 ? 32 : 48;
  p.qtiles = (B + TC_QM - 1) / TC_QM;
  p.dup = B <= 64 ? 1 : 0;
  // rows per tile: spread the DB over all SMs when it is small, 128-row tiles otherwise.  `sms` bounds the CTAs of the whole
  // launch: with several query tiles (grid = G x qtiles) every query tile gets sms / qtiles CTAs, unless the DB is so small
  // that one tile per CTA covers it (qstream below: grid = G, the query tiles stream through the CTA)
  auto split = [&](int ctas) {
    const int per_sm = (N + ctas - 1) / ctas;
    p.tile_n = std::min(TC_TILE_MAX, std::max(16, (per_sm + 15) / 16 * 16));
    p.tiles = (N + p.tile_n - 1) / p.tile_n;
    p.G = std::min(p.tiles, ctas);
    p.tiles_per_cta = (p.tiles + p.G - 1) / p.G;
  };
  split(sms);
  if (p.qtiles > 1 && p.tiles_per_cta > 1) split(std::max(1, sms / p.qtiles));
  if (p.tiles_per_cta > TC_MAX_TILES_PER_CTA) {  // keep the index field of the keys small: more CTAs than asked for
    p.tiles_per_cta = TC_MAX_TILES_PER_CTA;
    p.G = (p.tiles + TC_MAX_TILES_PER_CTA - 1) / TC_MAX_TILES_PER_CTA;
  }
  p.parts = (p.dup ? 2 : 1) * (scan_epi_warps() / 4);
  p.nsrc = p.parts * p.G;
  p.nb_bits = 7 + ceil_log2(p.tiles_per_cta);
  const int nch = D / 32;
  // several query tiles against a DB that gives every CTA one tile: keep the DB tile resident and stream the queries
  // (one launch wave, the DB is read once, no per-query-tile prologue)
  p.qstream = (p.qtiles > 1 && p.tiles_per_cta == 1) ? 1 : 0;
  p.a_tmem = p.qstream ? 0 : 1;
  if (p.qstream) {
    p.q_pitch = (int)align_up((size_t)p.tile_n * 128, 1024);  // resident operand: the DB tile
    p.st_pitch = TC_QM * 128;                                  // streamed operand: one 128-query chunk
    p.slack = 16384;  // the UMMA reads 128 rows from every chunk of the resident tile whatever its height
  } else {
    p.q_pitch = 0;    // the query tile lives in tensor memory
    p.st_pitch = p.tile_n * 128;
    p.slack = 0;
  }
  const size_t fixed = 1024 + (size_t)nch * p.q_pitch + p.slack + TC_CAND_BYTES + sizeof(ScanSmem) + 64;
  const size_t budget = 220 * 1024;
  if (fixed + 2 * (size_t)p.st_pitch > budget) return p;
  p.stages = (int)std::min<size_t>(TC_MAX_STAGES, (budget - fixed) / p.st_pitch);
  p.stages = std::min(p.stages, std::max(2, nch * (p.qstream ? p.qtiles : p.tiles_per_cta)));
  p.scan_smem = fixed + (size_t)p.stages * p.st_pitch;
  p.sel_smem = sizeof(SelSmem) + (size_t)D * 4 + (size_t)p.nsrc * p.KP * 4;
  if (p.sel_smem > 200 * 1024) return p;
  p.ok = true;
  return p;
}

size_t tc_workspace_bytes(const TcPlan& p, int B) {
  return align_up((size_t)B * p.nsrc * p.KP * sizeof(uint32_t), 256) + 256 /* norm bound */ + align_up((size_t)B * sizeof(int32_t), 256);
}

int launch_row_norm2_max(const float* db, int N, int D, float* out, cudaStream_t s) {
  T2P_CUDA(cudaMemsetAsync(out, 0, sizeof(float), s));
  const int blocks = std::max(1, std::min(592, (N + 7) / 8));
  row_norm2_max_kernel<<<blocks, 256, 0, s>>>(db, N, D, out);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

int launch_retrieve_tc(const TcPlan& p, const float* d_q, const float* d_db, int B, int N, int D, int k, int64_t idx_base,
                       const float* d_db_norm2_max, int force_rescan, double* d_out_scores, int64_t* d_out_idx, int32_t* d_stats,
                       void* d_ws, size_t ws_bytes, cudaStream_t s) {
  Arena a(d_ws, ws_bytes);
  uint32_t* part = a.take<uint32_t>((size_t)B * p.nsrc * p.KP);
  float* norm_slot = a.take<float>(1);
  int32_t* flagged = a.take<int32_t>((size_t)B);
  T2P_REQUIRE(a.ok, T2P_ERR_WORKSPACE, "retrieve_topk: workspace %zu < %zu bytes", ws_bytes, a.used);
  T2P_REQUIRE((reinterpret_cast<uintptr_t>(d_q) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_db) & 15) == 0, T2P_ERR_INVALID,
              "retrieve_topk: q and db must be 16-byte aligned for TMA");
  if (d_db_norm2_max == nullptr) {
    T2P_TRY(launch_row_norm2_max(d_db, N, D, norm_slot, s));
    d_db_norm2_max = norm_slot;
  }
  ScanParams sp;
  sp.B = B; sp.N = N; sp.D = D;
  sp.tile_n = p.tile_n;
  sp.q_box_rows = std::min(p.dup ? 64 : TC_QM, B);
  sp.db_box_rows = std::min(p.tile_n, N);
  sp.q_pitch = p.q_pitch; sp.st_pitch = p.st_pitch; sp.stages = p.stages;
  sp.nb_bits = p.nb_bits; sp.dup = p.dup; sp.qstream = p.qstream; sp.qtiles = p.qtiles;
  sp.a_tmem = p.a_tmem; sp.slack = p.slack;
#ifdef T2P_SCAN_TRACE  // knock-out experiments (wrong results by design) exist only in the trace build, never in the product library
  static const int scan_debug = [] { const char* e = getenv("T2P_SCAN_DEBUG"); return e ? atoi(e) : 0; }();
  sp.debug = scan_debug;
#else
  sp.debug = 0;
#endif
  sp.epi_warps = scan_epi_warps();
  CUtensorMap tq, tdb;
  T2P_TRY(make_tmap_rows(&tq, d_q, B, D, sp.q_box_rows));
  T2P_TRY(make_tmap_rows(&tdb, d_db, N, D, sp.db_box_rows));
  dim3 grid(p.G, p.qstream ? 1 : p.qtiles);
  // (launching the scan as 8-CTA clusters, so that it takes and returns SMs in the GPC-aligned groups the text encoder's
  // clusters need, was measured: 5-8 % slower per step than plain CTAs)
  if (p.KP == 16) {
    T2P_CUDA(cudaFuncSetAttribute(retrieve_scan_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.scan_smem));
    retrieve_scan_tc_kernel<16><<<grid, 32 * (2 + scan_epi_warps()), p.scan_smem, s>>>(tq, tdb, sp, d_q, part);
  } else {
    T2P_CUDA(cudaFuncSetAttribute(retrieve_scan_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.scan_smem));
    retrieve_scan_tc_kernel<32><<<grid, 32 * (2 + scan_epi_warps()), p.scan_smem, s>>>(tq, tdb, sp, d_q, part);
  }
  T2P_LAUNCH_CHECK();
  SelParams q;
  q.B = B; q.N = N; q.D = D; q.nsrc = p.nsrc; q.G = p.G; q.parts = p.parts; q.KP = p.KP; q.NC = p.NC; q.k = k;
  q.tile_n = p.tile_n; q.nb_bits = p.nb_bits; q.force_rescan = force_rescan; q.idx_base = idx_base;
  if (p.sel_smem > 48 * 1024)
    T2P_CUDA(cudaFuncSetAttribute(retrieve_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.sel_smem));
  // warp-per-query select first (the common case certifies there); the CTA-per-query kernel then only runs for flagged queries
  const size_t selw_smem = (size_t)SELW_WARPS * p.nsrc * p.KP * sizeof(uint32_t);
  const bool warp_path = !force_rescan && (D & 127) == 0 && D <= 256 && k <= 16 && p.nsrc <= 32 * SELW_MAX_LPL &&
                         selw_smem <= 200 * 1024 && flagged != nullptr;
  if (warp_path) {
    if (selw_smem > 48 * 1024)
      T2P_CUDA(cudaFuncSetAttribute(retrieve_select_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)selw_smem));
    retrieve_select_warp_kernel<<<(B + SELW_WARPS - 1) / SELW_WARPS, 32 * SELW_WARPS, selw_smem, s>>>(
        d_q, d_db, q, part, d_db_norm2_max, d_out_scores, d_out_idx, d_stats, flagged);
    T2P_LAUNCH_CHECK();
  }
  retrieve_select_kernel<<<warp_path ? std::min(B, 64) : B, SEL_THREADS, p.sel_smem, s>>>(
      d_q, d_db, q, part, d_db_norm2_max, d_out_scores, d_out_idx, d_stats, warp_path ? flagged : nullptr);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

}  // namespace t2p

#ifdef T2P_SCAN_TRACE
extern "C" int t2p_debug_scan_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, t2p::scan_trace, sizeof(t2p::scan_trace));
}
#endif
