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
//   warps 10-13 epilogue: tcgen05.ld 32 columns at a time, *2^-8 + bias, ReLU, transpose through shared memory, running
//               max over the rows of a centre (rows of a centre are contiguous), one atomicMax per (centre, column, tile).
// Two pipeline stages (A chunk 32 KB + W chunk 2*C*128 B each), two TMEM accumulators (epilogue of item i overlaps the
// MMAs of item i+1), two row tables.  No edge tensor in HBM, no scatter.
#include <cuda_fp16.h>

#include "kernels.h"
#include "sm100.cuh"

namespace t2p {

using namespace sm100;

constexpr int SAT_ROWS = 128;
constexpr int SAT_PROD_WARPS = 8;                       // A-operand producer warps (16 rows of a chunk each)
constexpr int SAT_EPI_WARP0 = 2 + SAT_PROD_WARPS;        // first of the 4 epilogue warps
constexpr int SAT_THREADS = 32 * (SAT_EPI_WARP0 + 4);
constexpr int SAT_STAGES = 2;
constexpr float SAT_WUNSCALE = 1.f / 256.f;
constexpr float SAT_AMAX = 60000.f;  // activations above this do not fit fp16 (max 65504): the layer is redone in fp32

struct SatRows {
  int rowT[SAT_ROWS];  // row of T (global point index) feeding edge r
  int rowS[SAT_ROWS];  // row of S / of the output (global centre index), -1 = padding
  int n_valid;         // 0 = nothing to do for this item
  int pad[3];
};

struct SatBars {
  uint64_t full[SAT_STAGES], empty[SAT_STAGES];
  uint64_t tmem_full[2], tmem_empty[2];
  uint64_t rows_full[2], rows_empty[2];
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
template <int K, int N, bool DENSE>
__global__ void __launch_bounds__(SAT_THREADS, 1)
sa_edge_tc_kernel(const float* __restrict__ T, const float* __restrict__ S, const int32_t* __restrict__ nbr,
                  const int32_t* __restrict__ cnt, const int32_t* __restrict__ obj_cell_start, int quirk, int P, int m,
                  int n_obj, const uint4* __restrict__ w_img, const float* __restrict__ b2,
                  float* __restrict__ out, int ldo, int32_t* __restrict__ overflow_flag) {
  constexpr int C = K;                           // row pitch of T / S
  constexpr int NKC = K / 64;                    // 64-wide K chunks
  constexpr int A_PART = SAT_ROWS * 128;         // one of {hi, lo} of an A chunk: 128 rows x 128 bytes
  constexpr int W_PART = N * 128;                // one of {hi, lo} of a W chunk: N rows x 128 bytes
  constexpr int STAGE_BYTES = 2 * A_PART + 2 * W_PART;
  constexpr int TMEM_COLS = 2 * N;               // two accumulators
  const int n_off = (int)blockIdx.y * N;         // first output column of this CTA
  w_img += (size_t)blockIdx.y * (NKC * 2 * W_PART / 16);
  b2 += n_off;
  out += n_off;
  constexpr int EPI_PITCH = 33;                  // floats per row of the transpose buffer (32 columns + 1: conflict-free)

  extern __shared__ __align__(1024) uint8_t sat_raw[];
  if ((smem_u32(sat_raw) & 1023u) != 0u) __trap();
  uint8_t* stages = sat_raw;
  float* epi = reinterpret_cast<float*>(stages + SAT_STAGES * STAGE_BYTES);  // [128][33]
  SatRows* rows = reinterpret_cast<SatRows*>(epi + SAT_ROWS * EPI_PITCH);    // [2]
  SatBars* bars = reinterpret_cast<SatBars*>(rows + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < SAT_STAGES; ++s) {
      mbar_init(&bars->full[s], SAT_PROD_WARPS + 1);  // the producer warps + the expect_tx arrive of the W loader
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 4);  // one lane per epilogue warp
      mbar_init(&bars->rows_full[a], 1);
      mbar_init(&bars->rows_empty[a], SAT_PROD_WARPS + 5);  // one lane of each consumer warp (MMA, producers, epilogue)
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(&bars->tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == 0) {
    // ===== row tables, one item ahead.  Objects are dealt round-robin to the CTAs; the per-centre edge counts of an object
    // are read once (the next object's are prefetched), its tiles are exactly ceil(E/128); a final table with
    // n_valid = -1 tells the consumers to stop. =====
    int it = 0;
    if (DENSE) {  // rows [128 tile, +128) of the n_obj input rows; output row = input row / m
      for (int tile = (int)blockIdx.x; tile * SAT_ROWS < n_obj; tile += (int)gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&bars->rows_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
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
    const int c0 = 2 * lane, c1 = 2 * lane + 1;
    int o = DENSE ? n_obj : (int)blockIdx.x;
    int n0 = 0, n1 = 0;
    if (o < n_obj) {
      n0 = c0 < m ? __ldg(cnt + (size_t)o * m + c0) + extra : 0;
      n1 = c1 < m ? __ldg(cnt + (size_t)o * m + c1) + extra : 0;
    }
    for (; o < n_obj; o += (int)gridDim.x) {
      const int o_next = o + (int)gridDim.x;
      int p0 = 0, p1 = 0;  // prefetch of the next object's counts
      if (o_next < n_obj) {
        p0 = c0 < m ? __ldg(cnt + (size_t)o_next * m + c0) + extra : 0;
        p1 = c1 < m ? __ldg(cnt + (size_t)o_next * m + c1) + extra : 0;
      }
      // inclusive prefix of the per-centre edge counts (m <= 64: two entries per lane)
      int inc = n0 + n1;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += v;
      }
      const int E = __shfl_sync(0xffffffffu, inc, 31);
      const int incl1 = inc, incl0 = inc - n1;  // inclusive prefix at c1, c0
      const int first = quirk ? __ldg(obj_cell_start + o) : 0;
      for (int e_base = 0; e_base < E; e_base += SAT_ROWS, ++it) {
        const int buf = it & 1;
        mbar_wait(&bars->rows_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
        SatRows* rw = rows + buf;
#pragma unroll
        for (int rr = 0; rr < SAT_ROWS / 32; ++rr) {
          const int r = rr * 32 + lane;
          const int e = e_base + r;
          int rt = 0, rs = -1;
          // smallest centre c with incl[c] > e: binary search over the 2m prefix values held two per lane.  Every lane
          // runs the same 6 rounds (m <= 64) -- the shuffles sit in convergent code -- and out-of-range rows are masked after.
          const int ee = min(e, E - 1);
          int lo = 0, hi = m - 1;
#pragma unroll
          for (int round = 0; round < 6; ++round) {
            const int mid = (lo + hi) >> 1;
            const int pv1 = __shfl_sync(0xffffffffu, incl1, mid >> 1), pv0 = __shfl_sync(0xffffffffu, incl0, mid >> 1);
            const int pv = (mid & 1) ? pv1 : pv0;
            if (lo < hi) {
              if (pv > ee) hi = mid; else lo = mid + 1;
            }
          }
          const int c = lo;
          const int q1 = __shfl_sync(0xffffffffu, incl1, c >> 1), q0 = __shfl_sync(0xffffffffu, incl0, c >> 1);
          const int nn1 = __shfl_sync(0xffffffffu, n1, c >> 1), nn0 = __shfl_sync(0xffffffffu, n0, c >> 1);
          if (e < E) {
            const int incl_c = (c & 1) ? q1 : q0;
            const int n_c = (c & 1) ? nn1 : nn0;
            const int slot = e - (incl_c - n_c);
            const int cn = n_c - extra;
            if (slot < cn) {
              rt = o * P + __ldg(nbr + ((size_t)o * m + c) * T2P_MAX_NEIGHBORS + slot);
            } else {  // flat-index self loop (see sa_edge_kernel)
              const int flat = (o - first) * m + c;
              rt = (first + flat / P) * P + flat % P;
            }
            rs = o * m + c;
          }
          rw->rowT[r] = rt;
          rw->rowS[r] = rs;
        }
        if (lane == 0) rw->n_valid = min(SAT_ROWS, E - e_base);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->rows_full[buf]);
      }
      n0 = p0;
      n1 = p1;
    }
    {  // terminator
      const int buf = it & 1;
      mbar_wait(&bars->rows_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
      if (lane == 0) rows[buf].n_valid = -1;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->rows_full[buf]);
    }
  } else if (warp == 1) {
    // ===== UMMA issuer =====
    const uint32_t idesc = sat_idesc(SAT_ROWS, N);
    const uint32_t st_addr = smem_u32(stages);
    int stage = 0, nv = 0;
    uint32_t ph = 0;
    for (int it = 0;; ++it) {
      const int buf = it & 1;
      mbar_wait(&bars->rows_full[buf], (uint32_t)((it >> 1) & 1));
      const int n_valid = rows[buf].n_valid;
      if (n_valid < 0) break;
      if (n_valid > 0) {
        const int acc = nv & 1;
        mbar_wait(&bars->tmem_empty[acc], (uint32_t)(((nv >> 1) & 1) ^ 1));
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * N;
        for (int kc = 0; kc < NKC; ++kc) {
          mbar_wait(&bars->full[stage], ph);
          tc_fence_after_sync();
          if (sat_elect_one()) {
            const uint32_t sa = st_addr + stage * STAGE_BYTES;
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {  // A_hi.W_hi, A_hi.W_lo, A_lo.W_hi
              const uint64_t a_desc = umma_desc_sw128_kmajor(sa + (prod == 2 ? A_PART : 0));
              const uint64_t b_desc = umma_desc_sw128_kmajor(sa + 2 * A_PART + (prod == 1 ? W_PART : 0));
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                sat_umma_ss(d_tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc, (kc | prod | ks) != 0);
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
      const int buf = it & 1;
      mbar_wait(&bars->rows_full[buf], (uint32_t)((it >> 1) & 1));
      const SatRows* rw = rows + buf;
      if (rw->n_valid < 0) break;
      if (rw->n_valid > 0) {
        for (int kc = 0; kc < NKC; ++kc) {
          mbar_wait(&bars->empty[stage], ph ^ 1);
          uint8_t* st = stages + stage * STAGE_BYTES;
          if (pw == 0 && lane == 0) {  // the W2 chunk of this K range: [hi | lo] images, contiguous in global memory
            mbar_expect_tx(&bars->full[stage], 2u * W_PART);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(w_img) + (size_t)kc * (2 * W_PART);
            const uint32_t dst = smem_u32(st + 2 * A_PART);
            for (uint32_t off = 0; off < 2u * W_PART; off += 16384u) sat_bulk_load(dst + off, src + off, 16384u, &bars->full[stage]);
          }
          const int col0 = kc * 64 + cg * 8;
          // all T loads of this lane (4 rows x 32 bytes) are issued before the first use: one L2 round trip per chunk
          // instead of one per row; the S rows repeat from row to row (a centre has up to 33 edges) and hit L1
          constexpr int RPL = SAT_ROWS / SAT_PROD_WARPS / 4;  // rows per lane and chunk
          float4 tv[RPL][2];
          int rsv[RPL];
#pragma unroll
          for (int i = 0; i < RPL; ++i) {
            const int r = pw * (SAT_ROWS / SAT_PROD_WARPS) + i * 4 + rsub;
            rsv[i] = rw->rowS[r];
            const float4* tp = reinterpret_cast<const float4*>(T + (size_t)rw->rowT[r] * C + col0);
            tv[i][0] = rsv[i] >= 0 ? __ldg(tp) : make_float4(0.f, 0.f, 0.f, 0.f);
            tv[i][1] = rsv[i] >= 0 ? __ldg(tp + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < RPL; ++i) {
            const int r = pw * (SAT_ROWS / SAT_PROD_WARPS) + i * 4 + rsub;
            const int rs = rsv[i];
            uint4 hi4 = make_uint4(0, 0, 0, 0), lo4 = make_uint4(0, 0, 0, 0);
            if (rs >= 0) {
              float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
              if (!DENSE) {
                const float4* sp = reinterpret_cast<const float4*>(S + (size_t)rs * C + col0);
                s0 = __ldg(sp);
                s1 = __ldg(sp + 1);
              }
              const float4 t0 = tv[i][0], t1 = tv[i][1];
              const float a[8] = {fmaxf(t0.x - s0.x, 0.f), fmaxf(t0.y - s0.y, 0.f), fmaxf(t0.z - s0.z, 0.f), fmaxf(t0.w - s0.w, 0.f),
                                  fmaxf(t1.x - s1.x, 0.f), fmaxf(t1.y - s1.y, 0.f), fmaxf(t1.z - s1.z, 0.f), fmaxf(t1.w - s1.w, 0.f)};
              uint32_t h[4], l[4];
#pragma unroll
              for (int j = 0; j < 8; ++j) amax = a[j] <= amax ? amax : a[j];  // NaN sticks (the comparison is false)
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
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->rows_empty[buf]);
    }
    // fp16 range guard: an activation beyond the fp16 range (or a NaN) makes this layer's tensor-core result unusable;
    // the flag makes the host-enqueued exact-fp32 kernel behind this launch recompute the layer (it is a no-op otherwise)
    if (!(amax <= SAT_AMAX) && overflow_flag) atomicOr(overflow_flag, 1);
  } else {
    // ===== epilogue: the last four warps own TMEM lane quadrants (warp & 3); 128 threads, named barrier 1 =====
    const int quad = warp & 3;
    const int row = quad * 32 + lane;          // TMEM lane = edge row of the tile
    const int et = (warp - SAT_EPI_WARP0) * 32 + lane;     // 0..127: thread index inside the epilogue group
    const int ccol = et & 31, rgrp = et >> 5;  // column pass: column of the 32-chunk, group of 32 rows
    int nv = 0;
    for (int it = 0;; ++it) {
      const int buf = it & 1;
      mbar_wait(&bars->rows_full[buf], (uint32_t)((it >> 1) & 1));
      const SatRows* rw = rows + buf;
      if (rw->n_valid < 0) break;
      if (rw->n_valid > 0) {
        const int acc = nv & 1;
        mbar_wait(&bars->tmem_full[acc], (uint32_t)((nv >> 1) & 1));
        tc_fence_after_sync();
        for (int cc = 0; cc < N / 32; ++cc) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * N + cc * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            epi[row * EPI_PITCH + j] = fmaxf(fmaf(__uint_as_float(v[j]), SAT_WUNSCALE, __ldg(b2 + cc * 32 + j)), 0.f);
          sat_named_barrier(1, 128);
          // running max over the rows of a centre (contiguous), one atomic per run
          {
            const int r0 = rgrp * 32;
            int cur = -1;
            float best = 0.f;
            for (int r = r0; r < r0 + 32; ++r) {
              const int rs = rw->rowS[r];
              if (rs != cur) {
                if (cur >= 0) atomic_max_nonneg(out + (size_t)cur * ldo + cc * 32 + ccol, best);
                cur = rs;
                best = 0.f;
              }
              if (rs >= 0) best = fmaxf(best, epi[r * EPI_PITCH + ccol]);
            }
            if (cur >= 0) atomic_max_nonneg(out + (size_t)cur * ldo + cc * 32 + ccol, best);
          }
          sat_named_barrier(1, 128);
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
static size_t sat_smem_bytes() {
  return (size_t)SAT_STAGES * (2 * SAT_ROWS * 128 + 2 * N * 128) + (size_t)SAT_ROWS * 33 * sizeof(float) + 2 * sizeof(SatRows) +
         sizeof(SatBars) + 64;
}

template <int C>
static int launch_sa_tc(const float* T, const float* S, const int32_t* nbr, const int32_t* cnt, const int32_t* obj_cell_start,
                        int quirk, int n_obj, int P, int m, const float* w_img, const float* b2, float* out, int sms,
                        int32_t* overflow_flag, cudaStream_t s) {
  const size_t smem = sat_smem_bytes<C>();
  T2P_REQUIRE(smem <= 227 * 1024, T2P_ERR_UNSUPPORTED, "set abstraction (tensor cores): %zu bytes of shared memory", smem);
  T2P_CUDA(cudaFuncSetAttribute(sa_edge_tc_kernel<C, C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(n_obj, sms);  // objects are dealt round-robin to persistent CTAs
  sa_edge_tc_kernel<C, C, false><<<grid, SAT_THREADS, smem, s>>>(T, S, nbr, cnt, obj_cell_start, quirk, P, m, n_obj,
                                                                reinterpret_cast<const uint4*>(w_img), b2, out, C, overflow_flag);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

bool sa_edge_tc_supported(int C1, int C2, int m) { return C1 == C2 && (C1 == 128 || C1 == 256) && m <= 64; }

int launch_sa_edge_tc(const float* T, const float* S, const int32_t* nbr, const int32_t* cnt, const int32_t* obj_cell_start,
                      int quirk, int n_obj, int P, int m, int C, const float* w_img, const float* b2, float* out, int sms,
                      int32_t* overflow_flag, cudaStream_t s) {
  if (n_obj <= 0) return T2P_OK;
  if (C == 128) return launch_sa_tc<128>(T, S, nbr, cnt, obj_cell_start, quirk, n_obj, P, m, w_img, b2, out, sms, overflow_flag, s);
  return launch_sa_tc<256>(T, S, nbr, cnt, obj_cell_start, quirk, n_obj, P, m, w_img, b2, out, sms, overflow_flag, s);
}

// y[M / group, N] = max over groups of `group` consecutive rows of relu(x[M, 512] . W + b): the second layer of the global
// abstraction MLP (512 -> 1024, pooled over the 32 points of an object).  x >= 0 is required (it is a ReLU output): the
// producers apply relu(x - 0).  `out` must be zero-filled.  N is processed in column blocks of 256 (blockIdx.y).
bool linear_groupmax_tc_supported(int K, int N) { return K == 512 && N % 256 == 0 && N >= 256; }

int launch_linear_groupmax_tc(const float* x, int M, int K, const float* w_img, const float* bias, int N, int group, float* out,
                              int sms, int32_t* overflow_flag, cudaStream_t s) {
  if (M <= 0) return T2P_OK;
  T2P_REQUIRE(linear_groupmax_tc_supported(K, N) && group >= 1, T2P_ERR_UNSUPPORTED, "linear_groupmax (tensor cores): K=%d N=%d", K, N);
  const size_t smem = sat_smem_bytes<256>();
  T2P_CUDA(cudaFuncSetAttribute(sa_edge_tc_kernel<512, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nblocks = N / 256, tiles = (M + SAT_ROWS - 1) / SAT_ROWS;
  dim3 grid(std::max(1, std::min(tiles, sms / nblocks)), nblocks);
  sa_edge_tc_kernel<512, 256, true><<<grid, SAT_THREADS, smem, s>>>(x, nullptr, nullptr, nullptr, nullptr, 0, 0, group, M,
                                                                   reinterpret_cast<const uint4*>(w_img), bias, out, N, overflow_flag);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

}  // namespace t2p
