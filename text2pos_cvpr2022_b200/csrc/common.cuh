// Shared host/device helpers for the text2pos_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/text2pos_b200.h"

// definition of the opaque handle of the C ABI
struct t2p_weights {
  float* d_blob;
  size_t n_floats;
};

namespace t2p {

void set_error(const char* fmt, ...);

// pointer to a packed sub-array; validates the range once per call site
static inline const float* wptr(const t2p_weights* w, int64_t off) { return off < 0 ? nullptr : w->d_blob + off; }
static inline bool lin_ok(const t2p_weights* w, const t2p_linear_desc& l) {
  if (l.k <= 0 || l.n <= 0 || l.w_off < 0) return false;
  if ((size_t)l.w_off + (size_t)l.k * l.n > w->n_floats) return false;
  if (l.b_off >= 0 && (size_t)l.b_off + l.n > w->n_floats) return false;
  return true;
}

#define T2P_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess) {                                                                       \
      ::t2p::set_error("%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);       \
      return T2P_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)

#define T2P_LAUNCH_CHECK() T2P_CUDA(cudaGetLastError())

#define T2P_REQUIRE(cond, code, ...)     \
  do {                                   \
    if (!(cond)) {                       \
      ::t2p::set_error(__VA_ARGS__);     \
      return (code);                     \
    }                                    \
  } while (0)

#define T2P_TRY(expr)             \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != T2P_OK) return rc__; \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller-provided workspace (256-byte aligned sub-buffers).
struct Arena {
  char* base;
  size_t size;
  size_t used;
  bool ok;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0), ok(true) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (base == nullptr || used + bytes > size) {
      ok = false;
      used += bytes;
      return nullptr;
    }
    T* r = reinterpret_cast<T*>(base + used);
    used += bytes;
    return r;
  }
};

static inline cudaStream_t as_stream(t2p_stream s) { return static_cast<cudaStream_t>(s); }

// ---- device helpers ------------------------------------------------------------------------------
#ifdef __CUDACC__
// squared distance with every product / sum individually rounded (no FMA contraction): the exact
// arithmetic the oracle uses (oracle/pointnet.py sqdist_f32), needed for bit-exact fps / ball query.
__device__ __forceinline__ float sqdist3_nofma(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  float d = __fmul_rn(dx, dx);
  d = __fadd_rn(d, __fmul_rn(dy, dy));
  d = __fadd_rn(d, __fmul_rn(dz, dz));
  return d;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// atomic max for NON-NEGATIVE floats (post-ReLU values): integer compare of the bit patterns
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}
#endif

}  // namespace t2p
