// Callers / data formats either side of the hot path, as device kernels:
//
//  * batch_object_points_kernel -- dataloading/kitti360pose/utils.py:89-110 + the transforms of
//    evaluation/pipeline.py:290-293 on a packed raw cell store: per object FixedPoints(P) (sampling WITH replacement,
//    counter-based indices or caller-given ones), NormalizeScale (centre on the mean, scale by 0.999999 / max|pos|), and
//    the object centre / mean colour over the RAW points (models/object_encoder.py:121-131 via
//    datapreparation/kitti360pose/imports.py:28-41).  One CTA per object.
//  * pose_head_kernel -- models/superglue_matcher.py:138-161 (get_pos_in_cell) for every (query, retrieved cell) pair:
//    mean over the matched objects of (object centre + offset of its hint), with and without the offsets, and the
//    confidence (number of matched objects, evaluation/pipeline.py:196).
//  * pose_accuracy_kernel -- evaluation/utils.py:31-54 (calc_sample_accuracies) + the mean-conf variant of
//    evaluation/pipeline.py:255-263: threshold hits per query, float64 like the numpy original.
#include "kernels.h"

namespace t2p {

// index i of object `obj` under seed: splitmix64 of the counter, scaled to [0, n) by the high 32 bits
__host__ __device__ __forceinline__ uint32_t fixed_points_index(uint64_t seed, uint64_t obj, uint32_t i, uint32_t P, uint32_t n) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (obj * (uint64_t)P + i + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(((z >> 32) * (uint64_t)n) >> 32);
}

constexpr int BOP_THREADS = 256;

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(BOP_THREADS)
batch_object_points_kernel(const float* __restrict__ raw_xyz, const float* __restrict__ raw_rgb,
                           const int64_t* __restrict__ obj_offsets, int P, const int32_t* __restrict__ choice, uint64_t seed,
                           int64_t obj_id_base, float* __restrict__ out_pos, float* __restrict__ out_rgb,
                           float* __restrict__ centers, float* __restrict__ mean_rgb, double* __restrict__ centers64,
                           int32_t* __restrict__ choice_out) {
  extern __shared__ __align__(16) float bop_smem[];  // [P][3] resampled positions
  __shared__ double red[6][BOP_THREADS / 32];
  __shared__ float s_mean[3];
  __shared__ float s_max[BOP_THREADS / 32];
  const int o = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t p0 = obj_offsets[o];
  const uint32_t n = (uint32_t)(obj_offsets[o + 1] - p0);
  const float* xyz = raw_xyz + 3 * p0;
  const float* rgb = raw_rgb + 3 * p0;

  // (1) centre / mean colour over the raw points, float64
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (uint32_t i = tid; i < n; i += BOP_THREADS) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      acc[c] += (double)__ldg(xyz + 3 * i + c);
      acc[3 + c] += (double)__ldg(rgb + 3 * i + c);
    }
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const double v = warp_sum_f64(acc[c]);
    if (lane == 0) red[c][warp] = v;
  }
  // (2) FixedPoints: resample with replacement
  for (int i = tid; i < P; i += BOP_THREADS) {
    const uint32_t j = choice ? (uint32_t)choice[(size_t)o * P + i] : fixed_points_index(seed, (uint64_t)(obj_id_base + o), i, P, n);
    if (choice_out) choice_out[(size_t)o * P + i] = (int32_t)j;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      bop_smem[3 * i + c] = __ldg(xyz + 3 * (size_t)j + c);
      out_rgb[((size_t)o * P + i) * 3 + c] = __ldg(rgb + 3 * (size_t)j + c);
    }
  }
  __syncthreads();
  if (tid < 6) {
    double s = 0;
    for (int w = 0; w < BOP_THREADS / 32; ++w) s += red[tid][w];
    s = n ? s / (double)n : 0.0;
    if (tid < 3) {
      centers[(size_t)o * 3 + tid] = (float)s;
      if (centers64) centers64[(size_t)o * 3 + tid] = s;
    } else {
      mean_rgb[(size_t)o * 3 + tid - 3] = (float)s;
    }
  }
  // (3) Center: the float32 mean accumulated row by row in index order (numpy's reduction order over axis 0, the order the
  // oracle's fixed_points_normalize sums in), one thread per coordinate
  if (tid >= 32 && tid < 35) {
    const int c = tid - 32;
    float s = 0.f;
    for (int i = 0; i < P; ++i) s = __fadd_rn(s, bop_smem[3 * i + c]);
    s_mean[c] = __fdiv_rn(s, (float)P);
  }
  __syncthreads();
  float mx = 0.f;
  for (int t = tid; t < 3 * P; t += BOP_THREADS) {
    const float v = __fsub_rn(bop_smem[t], s_mean[t % 3]);
    bop_smem[t] = v;
    mx = fmaxf(mx, fabsf(v));
  }
  mx = warp_max(mx);
  if (lane == 0) s_max[warp] = mx;
  __syncthreads();
  mx = s_max[0];
#pragma unroll
  for (int w = 1; w < BOP_THREADS / 32; ++w) mx = fmaxf(mx, s_max[w]);
  // (4) NormalizeScale: scale = float32((1 / max|pos|) * 0.999999), computed in float64 like the oracle
  const double m64 = (double)mx > 1e-12 ? (double)mx : 1e-12;
  const float scale = (float)(__dmul_rn(__ddiv_rn(1.0, m64), 0.999999));
  for (int t = tid; t < 3 * P; t += BOP_THREADS) out_pos[(size_t)o * P * 3 + t] = __fmul_rn(bop_smem[t], scale);
}

// ---------------------------------------------------------------------------------------------------------------
// pose head
// ---------------------------------------------------------------------------------------------------------------
// one thread per sample b = (query, retrieved cell): sequential float64 sums in object order, like np.mean over the list
__global__ void pose_head_kernel(const int64_t* __restrict__ matches0, int B, int M, int N, const float* __restrict__ offsets,
                                 const int64_t* __restrict__ off_idx, const double* __restrict__ centers,
                                 const int64_t* __restrict__ cell_idx, double* __restrict__ pos_mean, double* __restrict__ pos_off,
                                 int32_t* __restrict__ conf) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* ctr = centers + (size_t)(cell_idx ? cell_idx[b] : b) * M * 2;
  const float* off = offsets + (size_t)(off_idx ? off_idx[b] : b) * N * 2;
  double sx = 0, sy = 0, ox = 0, oy = 0;
  int cnt = 0;
  for (int i = 0; i < M; ++i) {
    const int64_t h = matches0[(size_t)b * M + i];
    if (h < 0 || h >= N) continue;
    const double cx = ctr[2 * i], cy = ctr[2 * i + 1];
    sx = __dadd_rn(sx, cx);
    sy = __dadd_rn(sy, cy);
    ox = __dadd_rn(ox, __dadd_rn(cx, (double)off[2 * h]));
    oy = __dadd_rn(oy, __dadd_rn(cy, (double)off[2 * h + 1]));
    ++cnt;
  }
  if (cnt > 0) {
    const double inv = (double)cnt;
    pos_mean[2 * b] = __ddiv_rn(sx, inv);
    pos_mean[2 * b + 1] = __ddiv_rn(sy, inv);
    pos_off[2 * b] = __ddiv_rn(ox, inv);
    pos_off[2 * b + 1] = __ddiv_rn(oy, inv);
  } else {  // no matches: the cell centre (models/superglue_matcher.py:159-160)
    pos_mean[2 * b] = pos_mean[2 * b + 1] = pos_off[2 * b] = pos_off[2 * b + 1] = 0.5;
  }
  conf[b] = cnt;
}

struct PoseAccParams {
  int32_t top_k[8];
  double threshs[8];
  int32_t n_k, n_t;
};

__device__ __forceinline__ double pose_dist(const double* pos, const double* origin, double size, double px, double py) {
  // pred_w = bbox_w[0:2] + pos * cell_size; ||pose_w - pred_w||_2 (numpy: sqrt(sum of squares), each op rounded)
  const double wx = __dadd_rn(origin[0], __dmul_rn(pos[0], size));
  const double wy = __dadd_rn(origin[1], __dmul_rn(pos[1], size));
  const double dx = __dsub_rn(px, wx), dy = __dsub_rn(py, wy);
  return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// one thread per query: hits[v][q][ik][it], v = 0 (mean), 1 (offsets), 2 (mean of the most confident cell; ik = 0 only)
__global__ void pose_accuracy_kernel(const double* __restrict__ pos_mean, const double* __restrict__ pos_off,
                                     const int32_t* __restrict__ conf, const int64_t* __restrict__ cell_idx, int Q, int K,
                                     const double* __restrict__ origin, const double* __restrict__ cell_size,
                                     const int32_t* __restrict__ cell_scene, const double* __restrict__ pose_w,
                                     const int32_t* __restrict__ pose_scene, const __grid_constant__ PoseAccParams p,
                                     int32_t* __restrict__ hits) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const double px = pose_w[2 * q], py = pose_w[2 * q + 1];
  const int ps = pose_scene[q];
  const size_t per = (size_t)Q * p.n_k * p.n_t;
  int best_c = 0, best_conf = -1;
  for (int v = 0; v < 2; ++v) {
    const double* pos = v ? pos_off : pos_mean;
    for (int ik = 0; ik < p.n_k; ++ik) {
      double mn = INFINITY;
      const int kk = min(p.top_k[ik], K);
      for (int c = 0; c < kk; ++c) {
        const int64_t ci = cell_idx[(size_t)q * K + c];
        double d = pose_dist(pos + 2 * ((size_t)q * K + c), origin + 2 * ci, cell_size[ci], px, py);
        if (cell_scene[ci] != ps) d = INFINITY;  // close-by cells of other scenes do not count
        mn = fmin(mn, d);
      }
      for (int it = 0; it < p.n_t; ++it) hits[v * per + ((size_t)q * p.n_k + ik) * p.n_t + it] = mn <= p.threshs[it] ? 1 : 0;
    }
  }
  for (int c = 0; c < K; ++c) {  // np.argmax: first maximum
    const int cf = conf[(size_t)q * K + c];
    if (cf > best_conf) { best_conf = cf; best_c = c; }
  }
  {
    const int64_t ci = cell_idx[(size_t)q * K + best_c];
    double d = pose_dist(pos_mean + 2 * ((size_t)q * K + best_c), origin + 2 * ci, cell_size[ci], px, py);
    if (cell_scene[ci] != ps) d = INFINITY;
    for (int ik = 0; ik < p.n_k; ++ik)
      for (int it = 0; it < p.n_t; ++it)
        hits[2 * per + ((size_t)q * p.n_k + ik) * p.n_t + it] = (ik == 0 && d <= p.threshs[it]) ? 1 : 0;
  }
}

}  // namespace t2p

using namespace t2p;

extern "C" {

uint32_t t2p_fixed_points_index(uint64_t seed, uint64_t obj, uint32_t i, uint32_t P, uint32_t n) {
  return fixed_points_index(seed, obj, i, P, n);
}

int t2p_batch_object_points(const float* d_raw_xyz, const float* d_raw_rgb, const int64_t* d_obj_offsets, int n_obj, int P,
                            const int32_t* d_choice, uint64_t seed, int64_t obj_id_base, float* d_pos, float* d_rgb,
                            float* d_centers, float* d_mean_rgb, double* d_centers64, int32_t* d_choice_out, t2p_stream stream) {
  T2P_REQUIRE(d_raw_xyz && d_raw_rgb && d_obj_offsets && d_pos && d_rgb && d_centers && d_mean_rgb, T2P_ERR_INVALID,
              "batch_object_points: null argument");
  T2P_REQUIRE(n_obj >= 0 && P >= 1 && P <= 4096, T2P_ERR_INVALID, "batch_object_points: n_obj=%d P=%d (1 <= P <= 4096)", n_obj, P);
  if (n_obj == 0) return T2P_OK;
  batch_object_points_kernel<<<n_obj, BOP_THREADS, (size_t)P * 3 * sizeof(float), as_stream(stream)>>>(
      d_raw_xyz, d_raw_rgb, d_obj_offsets, P, d_choice, seed, obj_id_base, d_pos, d_rgb, d_centers, d_mean_rgb, d_centers64,
      d_choice_out);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

int t2p_pose_head(const int64_t* d_matches0, int B, int M, int N, const float* d_offsets, const int64_t* d_off_idx,
                  const double* d_centers, const int64_t* d_cell_idx, double* d_pos_mean, double* d_pos_offsets,
                  int32_t* d_confidence, t2p_stream stream) {
  T2P_REQUIRE(d_matches0 && d_offsets && d_centers && d_pos_mean && d_pos_offsets && d_confidence, T2P_ERR_INVALID,
              "pose_head: null argument");
  T2P_REQUIRE(B >= 0 && M >= 1 && N >= 1, T2P_ERR_INVALID, "pose_head: B=%d M=%d N=%d", B, M, N);
  if (B == 0) return T2P_OK;
  pose_head_kernel<<<(B + 127) / 128, 128, 0, as_stream(stream)>>>(d_matches0, B, M, N, d_offsets, d_off_idx, d_centers,
                                                                   d_cell_idx, d_pos_mean, d_pos_offsets, d_confidence);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

int t2p_pose_accuracy(const double* d_pos_mean, const double* d_pos_offsets, const int32_t* d_confidence,
                      const int64_t* d_cell_idx, int Q, int K, const double* d_cell_origin, const double* d_cell_size,
                      const int32_t* d_cell_scene, const double* d_pose_w, const int32_t* d_pose_scene, const int32_t* h_top_k,
                      int n_k, const double* h_threshs, int n_t, int32_t* d_hits, t2p_stream stream) {
  T2P_REQUIRE(d_pos_mean && d_pos_offsets && d_confidence && d_cell_idx && d_cell_origin && d_cell_size && d_cell_scene &&
                  d_pose_w && d_pose_scene && h_top_k && h_threshs && d_hits,
              T2P_ERR_INVALID, "pose_accuracy: null argument");
  T2P_REQUIRE(Q >= 0 && K >= 1 && n_k >= 1 && n_k <= 8 && n_t >= 1 && n_t <= 8, T2P_ERR_INVALID,
              "pose_accuracy: Q=%d K=%d n_k=%d n_t=%d (at most 8 top-k values / thresholds)", Q, K, n_k, n_t);
  if (Q == 0) return T2P_OK;
  PoseAccParams p;
  p.n_k = n_k;
  p.n_t = n_t;
  for (int i = 0; i < n_k; ++i) p.top_k[i] = h_top_k[i];
  for (int i = 0; i < n_t; ++i) p.threshs[i] = h_threshs[i];
  pose_accuracy_kernel<<<(Q + 127) / 128, 128, 0, as_stream(stream)>>>(d_pos_mean, d_pos_offsets, d_confidence, d_cell_idx, Q, K,
                                                                       d_cell_origin, d_cell_size, d_cell_scene, d_pose_w,
                                                                       d_pose_scene, p, d_hits);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

}  // extern "C"
