// (a2) set abstraction, second local_nn layer + ReLU + max over the edge list on the tcgen05 tensor cores (C1 = C2 = C in
// {128, 256}: SA2 and SA3 of PointNet++, models/pointcloud/pointnet2.py:57-59 -- 85 % of the set-abstraction flops).
//
//   out[o, c, :] = max over the edges (j -> c) of object o of  relu( relu(T_j - S_c) . W2 + b2 )
//
// Persistent CTAs (objects dealt round-robin), one work item = 128 consecutive edges of one object (edges packed centre by
// centre; the packing, the
// self-loop quirk and the gather indices are those of the fp32 kernel sa_edge_kernel, csrc/dense.cu).  Fourteen warps:
//   warp 0      builds the row table of the NEXT item (prefix sum of the per-centre edge counts, one binary search per
//               row, neighbour lookup) while the current one is processed;
//   warp 1      issues the UMMAs: per 64-wide K chunk 4 K steps x 3 products of fp16 hi/lo splits (A_hi.W_hi + A_hi.W_lo +
//               A_lo.W_hi, fp32 accumulation in TMEM; the dropped lo.lo term is ~2^-22 relative: fp32-grade, 1e-4 target);
//   warps 2-9   produce the A operand: gather T_j (coalesced 32-byte pieces, 4 rows per warp instruction), subtract S_c,
//               ReLU, split into fp16 hi/lo, store into the 128-byte-swizzled K-major UMMA layout; warp 2 also streams the
//               W2 chunk (host-packed hi/lo images of 2^8.W2, cp.async.bulk) into the same stage;
//   warps 10-17 epilogue.  The product is TRANSPOSED -- D[channel][edge] = W2^T . A^T, i.e. the weight image is the UMMA "A"
//               operand (M = 128 output channels of this CTA's column block) and the activations the "B" operand (N = 128 edges) --
//               so a TMEM lane is an output channel and the columns are the edges: each epilogue thread owns one channel, reads
//               32 edges at a time with tcgen05.ld and keeps the running max over the (contiguous) edges of a centre in a register;
//               the centre boundaries are the same for every thread (warp-uniform control flow), bias + ReLU are applied once per
//               run (they commute with the max), and the 32 lanes of a warp store 32 consecutive channels (one coalesced atomicMax
//               per run).  No transpose through shared memory, no barrier.  (The row-major variant spent ~12,000 cycles per
//               128-edge item in the epilogue and bounded the kernel.)
// Two pipeline stages (A chunk 32 KB + W chunk 2*C*128 B each), two TMEM accumulators (epilogue of item i overlaps the
// MMAs of item i+1), two row tables.  No edge tensor in HBM, no scatter.
#include <cuda_fp16.h>

#include "kernels.h"
#include "sm100.cuh"

namespace t2p {

using namespace sm100;

constexpr int SAT_ROWS = 128;
constexpr int SAT_PROD_WARPS = 8;                       // A-operand producer warps (16 rows of a chunk each)
constexpr int SAT_EPI_WARP0 = 2 + SAT_PROD_WARPS;        // first of the 8 epilogue warps (two per TMEM lane quadrant)
constexpr int SAT_EPI_WARPS = 8;
constexpr int SAT_BUILDER2 = SAT_EPI_WARP0 + SAT_EPI_WARPS;  // second row-table warp (odd items)
constexpr int SAT_THREADS = 32 * (SAT_BUILDER2 + 1);
constexpr int SAT_MAX_STAGES = 4;
constexpr int SAT_NTAB = 6;   // row tables in flight: a table lives from the builder through producers, MMA and epilogue (~4 role
                              // latencies); with two tables at most two items were in the pipe and every role idled most of the time
constexpr float SAT_WUNSCALE = 1.f / 256.f;
constexpr float SAT_AMAX = 60000.f;  // activations above this do not fit fp16 (max 65504): the layer is redone in fp32

struct SatRows {
  int rowT[SAT_ROWS];  // row of T (global point index) feeding edge r
  int rowS[SAT_ROWS];  // row of S / of the output (global centre index), -1 = padding
  int n_valid;         // 0 = nothing to do for this item
  int t_row0;          // gathered mode: first row of T of this item's object (its tile in shared memory = rows [t_row0, t_row0 + P))
  int t_seq;           // sequence number of the object among this CTA's objects (parity of the tile barriers)
  int t_last;          // 1 = last item of its object (the producers release the tile after it)
};

struct SatBars {
  uint64_t full[SAT_MAX_STAGES], empty[SAT_MAX_STAGES];
  uint64_t w_full;  // resident-weight mode: the whole W2 image has landed
  uint64_t tmem_full[2], tmem_empty[2];
  uint64_t rows_full[SAT_NTAB], rows_empty[SAT_NTAB];
  uint64_t t_full, t_empty;  // gathered mode: the object's T tile has landed in shared memory / all producers are done with it
  uint32_t tmem_slot;
};

__device__ __forceinline__ void sat_umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ bool sat_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void sat_bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void sat_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sat_named_barrier(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__host__ __device__ constexpr uint32_t sat_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);  // fp16 x fp16 -> fp32, both K-major
}

// K = input channels, N = output channels of this launch's column block (blockIdx.y selects it), DENSE = false: set
// abstraction (gathered edges, K == N == C); DENSE = true: a plain linear layer + ReLU + max over groups of `m` consecutive
// rows (the global abstraction layer of PointNet++, models/pointcloud/pointnet2.py:45-49): rows = n_obj, T = the input.
// PLAIN (DENSE only): no max -- y[row, :] = act(x[row, :K] . W + b2 + pos[row, :3] . Wp) stored row by row: the first local_nn
// layers of the set abstractions (T_j = W1 . [x_j | pos_j] + b1: the x part on the tensor cores, the 3-wide position part in
// the epilogue), the first global-abstraction layer and lin1 / lin2.  x >= 0 is required as for DENSE (ReLU outputs / colours).
// Gathered mode (DENSE = false): CT = channels of T / S (row pitch); K = CT rounded up to 64 (zero columns), N = this launch's
// column block (blockIdx.y).  The T rows of ONE object (P x CT floats, contiguous) are staged in shared memory with one bulk
// copy per object and the edge gather reads them from there: an object's tile is touched by one CTA only, and gathering 128-512
// byte rows from L2 / HBM per edge (up to 33 edges per centre) was what bound the kernel (ncu: tensor pipe 12-24 %, 63 % of
// the samples on the long scoreboard).  Only the flat-index self-loop rows, which may belong to another object of the cell,
// still come from global memory.
// NST = pipeline stages; WRES: the whole fp16 hi/lo image of W2 (K/64 chunks x 2 x N x 128 bytes) stays resident in shared
// memory (loaded once per CTA) and the stages hold the A operand only -- re-streaming a W chunk per 128 edges put one
// L2 round trip (~1.5 us) on every chunk of a two-stage ring.
// TILE: stage the object's T rows in shared memory (else the producers gather every row from global memory / L2).
template <int K, int N, bool DENSE, bool PLAIN = false, int CT = K, int NST = 2, bool WRES = false, bool TILE = false>
__global__ void __launch_bounds__(SAT_THREADS, 1)
sa_edge_tc_kernel(const float* __restrict__ T, const float* __restrict__ S, const int32_t* __restrict__ nbr,
                  const int32_t* __restrict__ cnt, const int32_t* __restrict__ obj_cell_start, int quirk, int P, int m,
                  int n_obj, const uint4* __restrict__ w_img, const float* __restrict__ b2,
                  float* __restrict__ out, int ldo, int32_t* __restrict__ overflow_flag, int ldx = K,
                  const float* __restrict__ pos = nullptr, const float* __restrict__ Wp = nullptr, int relu_out = 1) {
  const int C = DENSE ? ldx : CT;                // row pitch of T / S
  constexpr int NKC = K / 64;                    // 64-wide K chunks
  constexpr int A_PART = SAT_ROWS * 128;         // one of {hi, lo} of an A chunk: 128 rows x 128 bytes
  constexpr int W_PART = N * 128;                // one of {hi, lo} of a W chunk: N rows x 128 bytes
  constexpr int STAGE_BYTES = 2 * A_PART + (WRES ? 0 : 2 * W_PART);
  constexpr int SAT_STAGES = NST;
  static_assert(NST >= 2 && NST <= SAT_MAX_STAGES, "2..4 stages");
  static_assert(N == 128, "the column block of a CTA = UMMA M = 128 output channels");
  constexpr int TMEM_COLS = 2 * SAT_ROWS;        // two accumulators of 128 columns (= edges) each
  const int n_off = (int)blockIdx.y * N;         // first output column of this CTA
  w_img += (size_t)blockIdx.y * (NKC * 2 * W_PART / 16);
  b2 += n_off;
  out += n_off;
  if (PLAIN && Wp != nullptr) Wp += n_off;
  constexpr int NEH = 2;                         // epilogue halves: 4 warps each, half h takes the 32-edge groups h, h + 2

  extern __shared__ __align__(1024) uint8_t sat_raw[];
  if ((smem_u32(sat_raw) & 1023u) != 0u) __trap();
  uint8_t* stages = sat_raw;
  SatRows* rows = reinterpret_cast<SatRows*>(stages + SAT_STAGES * STAGE_BYTES);  // [SAT_NTAB]
  SatBars* bars = reinterpret_cast<SatBars*>(rows + SAT_NTAB);
  int* inclS_all = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2 builders][128] edge-count prefixes
  float* Tsm = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(inclS_all) + 1024);  // TILE: [P][CT] tile of the current object
  // resident-weight mode: [K chunk][hi|lo][N rows x 128 bytes] behind the tile, 1024-byte aligned (128-byte swizzle atoms)
  uint8_t* Wres = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(Tsm) + (TILE ? (size_t)P * CT * 4 : 0) + 1023) & ~(uintptr_t)1023);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < SAT_STAGES; ++s) {
      mbar_init(&bars->full[s], SAT_PROD_WARPS + (WRES ? 0 : 1));  // the producer warps (+ the expect_tx arrive of the W loader)
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->w_full, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 4 * NEH);  // one lane per (active) epilogue warp
    }
    for (int a = 0; a < SAT_NTAB; ++a) {
      mbar_init(&bars->rows_full[a], 1);
      mbar_init(&bars->rows_empty[a], SAT_PROD_WARPS + 1 + 4 * NEH);  // one lane of each consumer warp (MMA, producers, epilogue)
    }
    mbar_init(&bars->t_full, 1);
    mbar_init(&bars->t_empty, SAT_PROD_WARPS);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(&bars->tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == 0 || warp == SAT_BUILDER2) {
    // Two row-table warps in gathered mode: warp 0 builds the even items, warp SAT_BUILDER2 the odd ones (a table costs ~7,500
    // cycles -- binary search + a dependent neighbour-list load per row -- and one warp could not keep the pipe fed); both walk
    // all objects and items.  DENSE mode: warp 0 alone.
    const int which = warp == 0 ? 0 : 1;
    int* inclS = inclS_all + which * 128;
    if (!(DENSE && which == 1)) {
    // ===== row tables, several items ahead.  Objects are dealt round-robin to the CTAs; the per-centre edge counts of an object
    // are read once (the next object's are prefetched), its tiles are exactly ceil(E/128); a final table with
    // n_valid = -1 tells the consumers to stop. =====
    int it = 0;
    if (DENSE) {  // rows [128 tile, +128) of the n_obj input rows; output row = input row / m
      for (int tile = (int)blockIdx.x; tile * SAT_ROWS < n_obj; tile += (int)gridDim.x, ++it) {
        const int buf = it % SAT_NTAB;
        mbar_wait(&bars->rows_empty[buf], (uint32_t)(((it / SAT_NTAB) & 1) ^ 1));
        SatRows* rw = rows + buf;
#pragma unroll
        for (int rr = 0; rr < SAT_ROWS / 32; ++rr) {
          const int r = rr * 32 + lane, e = tile * SAT_ROWS + r;
          rw->rowT[r] = e < n_obj ? e : 0;
          rw->rowS[r] = e < n_obj ? e / m : -1;
        }
        if (lane == 0) rw->n_valid = min(SAT_ROWS, n_obj - tile * SAT_ROWS);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->rows_full[buf]);
      }
    }
    const int extra = quirk ? 1 : 0;
    int oseq = 0;  // objects of this CTA staged so far
    // per-centre edge counts of an object, four centres per lane (m <= 128): c = 4 lane + i
    int o = DENSE ? n_obj : (int)blockIdx.x;
    int nc[4] = {0, 0, 0, 0};
    if (o < n_obj) {
#pragma unroll
      for (int i = 0; i < 4; ++i) nc[i] = 4 * lane + i < m ? __ldg(cnt + (size_t)o * m + 4 * lane + i) + extra : 0;
    }
    for (; o < n_obj; o += (int)gridDim.x) {
      const int o_next = o + (int)gridDim.x;
      int pn[4] = {0, 0, 0, 0};  // prefetch of the next object's counts
      if (o_next < n_obj) {
#pragma unroll
        for (int i = 0; i < 4; ++i) pn[i] = 4 * lane + i < m ? __ldg(cnt + (size_t)o_next * m + 4 * lane + i) + extra : 0;
      }
      // inclusive prefix of the per-centre edge counts -> shared memory (only this warp reads it)
      int inc = nc[0] + nc[1] + nc[2] + nc[3];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += v;
      }
      const int E = __shfl_sync(0xffffffffu, inc, 31);
      {
        int run = inc - (nc[0] + nc[1] + nc[2] + nc[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          run += nc[i];
          inclS[4 * lane + i] = run;
        }
      }
      __syncwarp();
      const int first = quirk ? __ldg(obj_cell_start + o) : 0;
      if (TILE && which == 0 && E > 0) {  // stage the object's T tile: wait until the producers have released the previous object's, then ONE bulk copy
        if (oseq > 0) mbar_wait(&bars->t_empty, (uint32_t)((oseq - 1) & 1));
        if (lane == 0) {
          const uint32_t bytes = (uint32_t)P * CT * 4u;
          mbar_expect_tx(&bars->t_full, bytes);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(T + (size_t)o * P * CT);
          for (uint32_t off = 0; off < bytes; off += 16384u)
            sat_bulk_load(smem_u32(Tsm) + off, src + off, min(16384u, bytes - off), &bars->t_full);
        }
        __syncwarp();
      }
      for (int e_base = 0; e_base < E; e_base += SAT_ROWS, ++it) {
        if ((it & 1) != which) continue;  // the other row-table warp's item
        const int buf = it % SAT_NTAB;
        mbar_wait(&bars->rows_empty[buf], (uint32_t)(((it / SAT_NTAB) & 1) ^ 1));
        SatRows* rw = rows + buf;
        {
          // Four rows per lane, in PHASES so that the loads of the four rows overlap (a branchy per-row body serialised a
          // binary search and a dependent global load per row: ~8,600 cycles per table): (1) the four binary searches over the
          // prefix in shared memory run interleaved, (2) the four neighbour-list loads are issued together (clamped addresses,
          // no branch), (3) the rows are written.
          constexpr int RPLB = SAT_ROWS / 32;
          int lo[RPLB], hi[RPLB], ee[RPLB];
#pragma unroll
          for (int rr = 0; rr < RPLB; ++rr) {
            ee[rr] = min(e_base + rr * 32 + lane, E - 1);
            lo[rr] = 0;
            hi[rr] = m - 1;
          }
#pragma unroll
          for (int round = 0; round < 7; ++round) {  // smallest centre c with incl[c] > e (m <= 128: 7 rounds)
            int pv[RPLB];
#pragma unroll
            for (int rr = 0; rr < RPLB; ++rr) pv[rr] = inclS[(lo[rr] + hi[rr]) >> 1];
#pragma unroll
            for (int rr = 0; rr < RPLB; ++rr) {
              const int mid = (lo[rr] + hi[rr]) >> 1;
              const bool go = lo[rr] < hi[rr];
              const bool left = pv[rr] > ee[rr];
              hi[rr] = (go && left) ? mid : hi[rr];
              lo[rr] = (go && !left) ? mid + 1 : lo[rr];
            }
          }
          int slot[RPLB], cn[RPLB], nb[RPLB];
#pragma unroll
          for (int rr = 0; rr < RPLB; ++rr) {
            const int c = lo[rr];
            const int incl_c = inclS[c];
            const int n_c = incl_c - (c ? inclS[c - 1] : 0);
            slot[rr] = ee[rr] - (incl_c - n_c);
            cn[rr] = n_c - extra;
          }
#pragma unroll
          for (int rr = 0; rr < RPLB; ++rr)  // clamped: a self-loop row reads a valid (unused) entry
            nb[rr] = __ldg(nbr + ((size_t)o * m + lo[rr]) * T2P_MAX_NEIGHBORS + min(max(slot[rr], 0), T2P_MAX_NEIGHBORS - 1));
#pragma unroll
          for (int rr = 0; rr < RPLB; ++rr) {
            const int r = rr * 32 + lane;
            const int c = lo[rr];
            const int flat = (o - first) * m + c;  // flat-index self loop (see sa_edge_kernel)
            const int self_row = (first + flat / P) * P + flat % P;
            const bool live = e_base + r < E;
            rw->rowT[r] = live ? (slot[rr] < cn[rr] ? o * P + nb[rr] : self_row) : 0;
            rw->rowS[r] = live ? o * m + c : -1;
          }
        }
        if (lane == 0) {
          rw->n_valid = min(SAT_ROWS, E - e_base);
          rw->t_row0 = o * P;
          rw->t_seq = oseq;
          rw->t_last = (e_base + SAT_ROWS >= E) ? 1 : 0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->rows_full[buf]);
      }
      if (E > 0) ++oseq;
      __syncwarp();  // inclS is rewritten for the next object
#pragma unroll
      for (int i = 0; i < 4; ++i) nc[i] = pn[i];
    }
    if (DENSE || (it & 1) == which) {  // terminator (published by the warp that owns this item index)
      const int buf = it % SAT_NTAB;
      mbar_wait(&bars->rows_empty[buf], (uint32_t)(((it / SAT_NTAB) & 1) ^ 1));
      if (lane == 0) rows[buf].n_valid = -1;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->rows_full[buf]);
    }
    }
  } else if (warp == 1) {
    // ===== UMMA issuer =====
    const uint32_t idesc = sat_idesc(N, SAT_ROWS);  // M = output channels, N = edges
    const uint32_t st_addr = smem_u32(stages);
    if (WRES) {  // the whole W2 image, once
      if (lane == 0) {
        constexpr uint32_t WB = (uint32_t)NKC * 2u * W_PART;
        mbar_expect_tx(&bars->w_full, WB);
        for (uint32_t off = 0; off < WB; off += 16384u)
          sat_bulk_load(smem_u32(Wres) + off, reinterpret_cast<const uint8_t*>(w_img) + off, min(16384u, WB - off), &bars->w_full);
      }
      mbar_wait(&bars->w_full, 0);
    }
    int stage = 0, nv = 0;
    uint32_t ph = 0;
    for (int it = 0;; ++it) {
      const int buf = it % SAT_NTAB;
      mbar_wait(&bars->rows_full[buf], (uint32_t)((it / SAT_NTAB) & 1));
      const int n_valid = rows[buf].n_valid;
      if (n_valid < 0) break;
      if (n_valid > 0) {
        const int acc = nv & 1;
        mbar_wait(&bars->tmem_empty[acc], (uint32_t)(((nv >> 1) & 1) ^ 1));
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * SAT_ROWS;
        for (int kc = 0; kc < NKC; ++kc) {
          mbar_wait(&bars->full[stage], ph);
          tc_fence_after_sync();
          if (sat_elect_one()) {
            const uint32_t sa = st_addr + stage * STAGE_BYTES;
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {  // W_hi.A_hi, W_lo.A_hi, W_hi.A_lo  (transposed product: W is the "A" operand)
              const uint64_t act_desc = umma_desc_sw128_kmajor(sa + (prod == 2 ? A_PART : 0));
              const uint64_t w_desc = umma_desc_sw128_kmajor((WRES ? smem_u32(Wres) + (uint32_t)kc * 2u * W_PART : sa + 2 * A_PART) +
                                                             (prod == 1 ? W_PART : 0));
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                sat_umma_ss(d_tmem, w_desc + 2 * ks, act_desc + 2 * ks, idesc, (kc | prod | ks) != 0);
            }
            umma_commit(&bars->empty[stage]);
            if (kc == NKC - 1) umma_commit(&bars->tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == SAT_STAGES) { stage = 0; ph ^= 1; }
        }
        ++nv;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->rows_empty[buf]);
    }
  } else if (warp < SAT_EPI_WARP0) {
    // ===== A producers: warp pw handles rows 16 pw .. 16 pw + 15; lane = (row sub-index 0..3, 8-column group 0..7) =====
    const int pw = warp - 2;
    const int cg = lane & 7, rsub = lane >> 3;
    int stage = 0;
    uint32_t ph = 0;
    float amax = 0.f;  // largest activation converted to fp16 by this lane (range guard, see the end of this branch)
    for (int it = 0;; ++it) {
      const int buf = it % SAT_NTAB;
      mbar_wait(&bars->rows_full[buf], (uint32_t)((it / SAT_NTAB) & 1));
      const SatRows* rw = rows + buf;
      if (rw->n_valid < 0) break;
      if (rw->n_valid > 0) {
        const int t_row0 = DENSE ? 0 : rw->t_row0;
        if (TILE) {  // the tile of this item's object (a completed phase stays complete: later items of the object pass at once)
          mbar_wait(&bars->t_full, (uint32_t)(rw->t_seq & 1));
        }
        for (int kc = 0; kc < NKC; ++kc) {
          mbar_wait(&bars->empty[stage], ph ^ 1);
          uint8_t* st = stages + stage * STAGE_BYTES;
          if (!WRES && pw == 0 && lane == 0) {  // the W2 chunk of this K range: [hi | lo] images, contiguous in global memory
            mbar_expect_tx(&bars->full[stage], 2u * W_PART);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(w_img) + (size_t)kc * (2 * W_PART);
            const uint32_t dst = smem_u32(st + 2 * A_PART);
            for (uint32_t off = 0; off < 2u * W_PART; off += 16384u) sat_bulk_load(dst + off, src + off, 16384u, &bars->full[stage]);
          }
          const int col0 = kc * 64 + cg * 8;
          // all T loads of this lane (4 rows x 32 bytes) are issued before the first use: one L2 round trip per chunk
          // instead of one per row; the S rows repeat from row to row (a centre has up to 33 edges) and hit L1
          constexpr int RPL = SAT_ROWS / SAT_PROD_WARPS / 4;  // rows per lane and chunk
          float4 tv[RPL][2], sv[RPL][2];
          int rsv[RPL];
#pragma unroll
          for (int i = 0; i < RPL; ++i) {
            const int r = pw * (SAT_ROWS / SAT_PROD_WARPS) + i * 4 + rsub;
            rsv[i] = rw->rowS[r];
            const int rt = rw->rowT[r];
            const bool live = rsv[i] >= 0 && (DENSE || col0 < CT);  // columns >= CT are the zero padding of K
            if (TILE && (unsigned)(rt - t_row0) < (unsigned)P) {    // the usual case: a row of this object's tile
              const float4* tp = reinterpret_cast<const float4*>(Tsm + (size_t)(rt - t_row0) * CT + col0);
              tv[i][0] = live ? tp[0] : make_float4(0.f, 0.f, 0.f, 0.f);
              tv[i][1] = live ? tp[1] : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
              const float4* tp = reinterpret_cast<const float4*>(T + (size_t)rt * C + col0);
              tv[i][0] = live ? __ldg(tp) : make_float4(0.f, 0.f, 0.f, 0.f);
              tv[i][1] = live ? __ldg(tp + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            sv[i][0] = sv[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!DENSE && live) {  // the centre's S row: issued with the T loads (one memory round trip per chunk, not two)
              const float4* sp = reinterpret_cast<const float4*>(S + (size_t)rsv[i] * C + col0);
              sv[i][0] = __ldg(sp);
              sv[i][1] = __ldg(sp + 1);
            }
          }
#pragma unroll
          for (int i = 0; i < RPL; ++i) {
            const int r = pw * (SAT_ROWS / SAT_PROD_WARPS) + i * 4 + rsub;
            const int rs = rsv[i];
            uint4 hi4 = make_uint4(0, 0, 0, 0), lo4 = make_uint4(0, 0, 0, 0);
            if (rs >= 0 && (DENSE || col0 < CT)) {
              const float4 s0 = sv[i][0], s1 = sv[i][1];
              const float4 t0 = tv[i][0], t1 = tv[i][1];
              const float a[8] = {fmaxf(t0.x - s0.x, 0.f), fmaxf(t0.y - s0.y, 0.f), fmaxf(t0.z - s0.z, 0.f), fmaxf(t0.w - s0.w, 0.f),
                                  fmaxf(t1.x - s1.x, 0.f), fmaxf(t1.y - s1.y, 0.f), fmaxf(t1.z - s1.z, 0.f), fmaxf(t1.w - s1.w, 0.f)};
              uint32_t h[4], l[4];
#pragma unroll
              for (int j = 0; j < 8; ++j) amax = fmaxf(amax, a[j]);  // (a >= 0 after the ReLU; one instruction per value)
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
            const uint32_t off = (uint32_t)r * 128u + (uint32_t)((cg ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(st + off) = hi4;
            *reinterpret_cast<uint4*>(st + A_PART + off) = lo4;
          }
          sat_fence_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->full[stage]);
          if (++stage == SAT_STAGES) { stage = 0; ph ^= 1; }
        }
        if (TILE && rw->t_last) {  // this warp has read everything it needs from the tile
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->t_empty);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->rows_empty[buf]);
    }
    // fp16 range guard: an activation beyond the fp16 range (or a NaN) makes this layer's tensor-core result unusable;
    // the flag makes the host-enqueued exact-fp32 kernel behind this launch recompute the layer (it is a no-op otherwise)
    if (!(amax <= SAT_AMAX) && overflow_flag) atomicOr(overflow_flag, 1);
  } else {
    // ===== epilogue (transposed product): eight warps, two per TMEM lane quadrant (warp & 3); lane = output channel; half h
    // takes the 32-edge column groups cc = h, h + 2 of every item =====
    const int quad = warp & 3;
    const int eh = (warp - SAT_EPI_WARP0) >> 2;
    const int ch = quad * 32 + lane;             // channel inside this CTA's column block
    const bool ch_ok = n_off + ch < ldo;         // (sa1: the block is padded from 64 to 128 channels)
    const float bb = ch_ok ? __ldg(b2 + ch) : 0.f;
    float wp0 = 0.f, wp1 = 0.f, wp2 = 0.f;
    if (PLAIN && Wp != nullptr && ch_ok) {
      wp0 = __ldg(Wp + ch);
      wp1 = __ldg(Wp + ldo + ch);
      wp2 = __ldg(Wp + 2 * ldo + ch);
    }
    int nv = 0;
    for (int it = 0;; ++it) {
      const int buf = it % SAT_NTAB;
      mbar_wait(&bars->rows_full[buf], (uint32_t)((it / SAT_NTAB) & 1));
      const SatRows* rw = rows + buf;
      if (rw->n_valid < 0) break;
      if (rw->n_valid > 0) {
        const int acc = nv & 1;
        mbar_wait(&bars->tmem_full[acc], (uint32_t)((nv >> 1) & 1));
        tc_fence_after_sync();
        for (int cc = eh; cc < SAT_ROWS / 32; cc += NEH) {
          if (cc * 32 >= rw->n_valid || n_off + quad * 32 >= ldo) break;  // warp-uniform: no edge in this group / padded channels only
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * SAT_ROWS + cc * 32, v);
          int rsv[32];  // the centres (output rows) of these 32 edges: the same for every thread
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const int4 t4 = *reinterpret_cast<const int4*>(&rw->rowS[cc * 32 + 4 * j4]);
            rsv[4 * j4] = t4.x; rsv[4 * j4 + 1] = t4.y; rsv[4 * j4 + 2] = t4.z; rsv[4 * j4 + 3] = t4.w;
          }
          tmem_ld_wait();
          if (PLAIN) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int rs = rsv[j];
              if (rs >= 0) {  // warp-uniform
                float y = fmaf(__uint_as_float(v[j]), SAT_WUNSCALE, bb);
                if (pos != nullptr) {
                  const float* pp = pos + (size_t)rs * 3;
                  y = fmaf(__ldg(pp), wp0, y);
                  y = fmaf(__ldg(pp + 1), wp1, y);
                  y = fmaf(__ldg(pp + 2), wp2, y);
                }
                if (ch_ok) out[(size_t)rs * ldo + ch] = relu_out ? fmaxf(y, 0.f) : y;
              }
            }
          } else {
            // running max over the edges of a centre (contiguous columns), one atomic per run: relu(max(a) * 2^-8 + b)
            int cur = -1;
            float best = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int rs = rsv[j];
              if (rs != cur) {  // warp-uniform
                if (cur >= 0 && ch_ok) atomic_max_nonneg(out + (size_t)cur * ldo + ch, fmaxf(fmaf(best, SAT_WUNSCALE, bb), 0.f));
                cur = rs;
                best = -INFINITY;
              }
              if (rs >= 0) best = fmaxf(best, __uint_as_float(v[j]));
            }
            if (cur >= 0 && ch_ok) atomic_max_nonneg(out + (size_t)cur * ldo + ch, fmaxf(fmaf(best, SAT_WUNSCALE, bb), 0.f));
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
        ++nv;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->rows_empty[buf]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int N>
static size_t sat_smem_bytes(size_t tile_bytes = 0, int stages = 2, size_t wres_bytes = 0) {
  return (size_t)stages * (2 * SAT_ROWS * 128 + (wres_bytes ? 0 : 2 * N * 128)) + SAT_NTAB * sizeof(SatRows) + sizeof(SatBars) + 256 + 1024 + tile_bytes + (wres_bytes ? wres_bytes + 1024 : 0) + 64;
}

// CT = channels of the layer (T / S pitch, output width), K = CT padded to 64, NB = column block per CTA (grid.y = CT2 / NB)
template <int K, int NB, int CT, int NST, bool WRES, bool TILE>
static int launch_sa_tc(const float* T, const float* S, const int32_t* nbr, const int32_t* cnt, const int32_t* obj_cell_start,
                        int quirk, int n_obj, int P, int m, int C2, const float* w_img, const float* b2, float* out, int sms,
                        int32_t* overflow_flag, cudaStream_t s) {
  const size_t smem = sat_smem_bytes<NB>(TILE ? (size_t)P * CT * sizeof(float) : 0, NST, WRES ? (size_t)(K / 64) * 2 * NB * 128 : 0);
  T2P_REQUIRE(smem <= 227 * 1024, T2P_ERR_UNSUPPORTED, "set abstraction (tensor cores): %zu bytes of shared memory", smem);
  auto kern = sa_edge_tc_kernel<K, NB, false, false, CT, NST, WRES, TILE>;
  T2P_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nblocks = (C2 + NB - 1) / NB;
  dim3 grid(std::max(1, std::min(n_obj, sms / nblocks)), nblocks);  // objects are dealt round-robin to persistent CTAs
  kern<<<grid, SAT_THREADS, smem, s>>>(T, S, nbr, cnt, obj_cell_start, quirk, P, m, n_obj, reinterpret_cast<const uint4*>(w_img), b2,
                                       out, C2, overflow_flag, CT, nullptr, nullptr, 1);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

bool sa_edge_tc_supported(int C1, int C2, int m) {
  return m <= 128 && ((C1 == 32 && C2 == 64) || (C1 == 128 && C2 == 128) || (C1 == 256 && C2 == 256));
}

int launch_sa_edge_tc(const float* T, const float* S, const int32_t* nbr, const int32_t* cnt, const int32_t* obj_cell_start,
                      int quirk, int n_obj, int P, int m, int C, const float* w_img, const float* b2, float* out, int sms,
                      int32_t* overflow_flag, cudaStream_t s) {
  if (n_obj <= 0) return T2P_OK;
  if (C == 32)  // SA1: 32 -> 64, K padded to one 64-wide chunk
    return launch_sa_tc<64, 128, 32, 4, true, true>(T, S, nbr, cnt, obj_cell_start, quirk, n_obj, P, m, 64, w_img, b2, out, sms, overflow_flag, s);
  if (C == 128) return launch_sa_tc<128, 128, 128, 4, true, false>(T, S, nbr, cnt, obj_cell_start, quirk, n_obj, P, m, 128, w_img, b2, out, sms, overflow_flag, s);
  // SA3: two 128-column blocks per object (one CTA each), W2 streamed chunk by chunk (the 128 KB image of a block is not resident)
  return launch_sa_tc<256, 128, 256, 3, false, false>(T, S, nbr, cnt, obj_cell_start, quirk, n_obj, P, m, 256, w_img, b2, out, sms, overflow_flag, s);
}

// y[M / group, N] = max over groups of `group` consecutive rows of relu(x[M, 512] . W + b): the second layer of the global
// abstraction MLP (512 -> 1024, pooled over the 32 points of an object).  x >= 0 is required (it is a ReLU output): the
// producers apply relu(x - 0).  `out` must be zero-filled.  N is processed in column blocks of 256 (blockIdx.y).
bool linear_groupmax_tc_supported(int K, int N) { return K == 512 && N % 128 == 0 && N >= 128; }

int launch_linear_groupmax_tc(const float* x, int M, int K, const float* w_img, const float* bias, int N, int group, float* out,
                              int sms, int32_t* overflow_flag, cudaStream_t s) {
  if (M <= 0) return T2P_OK;
  T2P_REQUIRE(linear_groupmax_tc_supported(K, N) && group >= 1, T2P_ERR_UNSUPPORTED, "linear_groupmax (tensor cores): K=%d N=%d", K, N);
  const size_t smem = sat_smem_bytes<128>(0, 3);
  auto kern = sa_edge_tc_kernel<512, 128, true, false, 512, 3, false, false>;
  T2P_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nblocks = N / 128, tiles = (M + SAT_ROWS - 1) / SAT_ROWS;
  dim3 grid(std::max(1, std::min(tiles, sms / nblocks)), nblocks);
  kern<<<grid, SAT_THREADS, smem, s>>>(x, nullptr, nullptr, nullptr, nullptr, 0, 0, group, M,
                                       reinterpret_cast<const uint4*>(w_img), bias, out, N, overflow_flag, 512, nullptr, nullptr, 1);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

// y[M, N] = act(x[M, K] (ld ldx, x >= 0) . W + bias + pos[M, 3] . Wp): dense layer on the tensor cores, N in column blocks of NB
template <int K>
static int launch_linear_tc_t(const float* x, int ldx, int M, const float* w_img, const float* bias, int N, const float* pos,
                              const float* Wp, bool relu, float* y, int sms, int32_t* overflow_flag, cudaStream_t s) {
  constexpr int NB = 128;
  const size_t smem = sat_smem_bytes<NB>(0, 3);
  auto kern = sa_edge_tc_kernel<K, NB, true, true, K, 3, false, false>;
  T2P_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nblocks = N / NB, tiles = (M + SAT_ROWS - 1) / SAT_ROWS;
  dim3 grid(std::max(1, std::min(tiles, sms / nblocks)), nblocks);
  kern<<<grid, SAT_THREADS, smem, s>>>(x, nullptr, nullptr, nullptr, nullptr, 0, 0, 1, M, reinterpret_cast<const uint4*>(w_img), bias,
                                       y, N, overflow_flag, ldx, pos, Wp, relu ? 1 : 0);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

bool linear_tc_supported(int K, int N) {
  return (K == 64 && N == 128) || (K == 128 && N == 256) || (K == 256 && N == 512) || (K == 1024 && N == 512) || (K == 512 && N == 256);
}

int launch_linear_tc(const float* x, int ldx, int M, int K, const float* w_img, const float* bias, int N, const float* pos,
                     const float* Wp, bool relu, float* y, int sms, int32_t* overflow_flag, cudaStream_t s) {
  if (M <= 0) return T2P_OK;
  T2P_REQUIRE(linear_tc_supported(K, N), T2P_ERR_UNSUPPORTED, "linear (tensor cores): K=%d N=%d", K, N);
  if (K == 64) return launch_linear_tc_t<64>(x, ldx, M, w_img, bias, N, pos, Wp, relu, y, sms, overflow_flag, s);
  if (K == 128) return launch_linear_tc_t<128>(x, ldx, M, w_img, bias, N, pos, Wp, relu, y, sms, overflow_flag, s);
  if (K == 256) return launch_linear_tc_t<256>(x, ldx, M, w_img, bias, N, pos, Wp, relu, y, sms, overflow_flag, s);
  if (K == 1024) return launch_linear_tc_t<1024>(x, ldx, M, w_img, bias, N, pos, Wp, relu, y, sms, overflow_flag, s);
  return launch_linear_tc_t<512>(x, ldx, M, w_img, bias, N, pos, Wp, relu, y, sms, overflow_flag, s);
}

}  // namespace t2p
