// Warp-distributed sorted top-k list: lane i (< kp <= 32) holds the i-th best (score, index) entry.
// Ordering = (score descending, index ascending), the oracle's stable-sort rule (oracle/retrieval.py).
#pragma once
#include "common.cuh"

namespace t2p {

template <typename S, typename I>
struct WarpTopK {
  S s;   // this lane's score
  I i;   // this lane's index
  int kp;

  __device__ __forceinline__ static bool better(S sa, I ia, S sb, I ib) { return sa > sb || (sa == sb && ia < ib); }

  __device__ __forceinline__ void init(int kp_, S neg_inf, I max_idx) {
    kp = kp_;
    s = neg_inf;
    i = max_idx;
  }
  // warp-uniform candidate; every lane calls
  __device__ __forceinline__ void insert(S cs, I ci) {
    const int lane = threadIdx.x & 31;
    const bool mine_better = (lane < kp) && !better(cs, ci, s, i);  // equal entries count as "already there"
    const int pos = __popc(__ballot_sync(0xffffffffu, mine_better));
    const S ps = __shfl_up_sync(0xffffffffu, s, 1);
    const I pi = __shfl_up_sync(0xffffffffu, i, 1);
    if (pos < kp) {
      if (lane > pos) {
        s = ps;
        i = pi;
      } else if (lane == pos) {
        s = cs;
        i = ci;
      }
    }
  }
  // per-lane candidates (valid flag per lane); inserts all that beat the current k-th entry
  __device__ __forceinline__ void offer(bool valid, S cs, I ci) {
    S ts = __shfl_sync(0xffffffffu, s, kp - 1);
    I ti = __shfl_sync(0xffffffffu, i, kp - 1);
    unsigned mask = __ballot_sync(0xffffffffu, valid && better(cs, ci, ts, ti));
    while (mask) {
      const int b = __ffs(mask) - 1;
      mask &= mask - 1;
      const S bs = __shfl_sync(0xffffffffu, cs, b);
      const I bi = __shfl_sync(0xffffffffu, ci, b);
      insert(bs, bi);
    }
  }
};

__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }

}  // namespace t2p
