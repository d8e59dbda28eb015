// Index kernels of the object/cell encoders: farthest point sampling + ball query (one warp per object,
// points resident in shared memory / registers) and the per-cell k-nearest-neighbour search of
// DynamicEdgeConv.  All distance arithmetic is the oracle's (individually rounded, no FMA) so that the
// produced indices are bit-exact.
#include "kernels.h"

namespace t2p {

constexpr int FB_WARPS = 4;

// mode 0: run FPS, write ctr_idx (+cpos);  mode 1: read centres from ctr_idx.  do_ball: also run the ball query.
template <int PPL>
__global__ void __launch_bounds__(FB_WARPS * 32)
fps_ball_kernel(const float* __restrict__ pos, int n_obj, int P, int m, float r2, int mode, int do_ball,
                int32_t* __restrict__ ctr_idx, float* __restrict__ cpos, int32_t* __restrict__ nbr,
                int32_t* __restrict__ cnt) {
  extern __shared__ float fb_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * FB_WARPS + warp;
  if (o >= n_obj) return;  // warp-uniform; only __syncwarp below
  float* px = fb_smem + (size_t)warp * (3 * P + m);
  float* py = px + P;
  float* pz = py + P;
  int* sidx = reinterpret_cast<int*>(pz + P);

  const float* src = pos + (size_t)o * P * 3;
  for (int t = lane; t < 3 * P; t += 32) {
    const float v = __ldg(src + t);
    const int j = t / 3, c = t - 3 * j;
    (c == 0 ? px : (c == 1 ? py : pz))[j] = v;
  }
  __syncwarp();

  if (mode == 0) {
    float x[PPL], y[PPL], z[PPL], mind[PPL];
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
      const int j = lane + 32 * i;
      const bool valid = j < P;
      x[i] = valid ? px[j] : 0.f;
      y[i] = valid ? py[j] : 0.f;
      z[i] = valid ? pz[j] : 0.f;
      mind[i] = valid ? __int_as_float(0x7f800000) : -1.f;
    }
    int last = 0;
    if (lane == 0) sidx[0] = 0;
    for (int s = 1; s < m; ++s) {
      const float lx = px[last], ly = py[last], lz = pz[last];
      float best = -2.f;
      int bi = 0x7fffffff;
#pragma unroll
      for (int i = 0; i < PPL; ++i) {
        const float d = sqdist3_nofma(x[i], y[i], z[i], lx, ly, lz);
        mind[i] = fminf(mind[i], d);
        if (mind[i] > best) {  // ascending i == ascending index: strict '>' keeps the lowest index
          best = mind[i];
          bi = lane + 32 * i;
        }
      }
      const int bb = __float_as_int(best);  // non-negative floats order like signed ints; -1/-2 are negative ints
      const int mx = __reduce_max_sync(0xffffffffu, bb);
      const unsigned cand = (bb == mx) ? (unsigned)bi : 0xffffffffu;
      last = (int)__reduce_min_sync(0xffffffffu, cand);
      if (lane == 0) sidx[s] = last;
    }
    __syncwarp();
    for (int c = lane; c < m; c += 32) ctr_idx[(size_t)o * m + c] = sidx[c];
  } else {
    for (int c = lane; c < m; c += 32) sidx[c] = ctr_idx[(size_t)o * m + c];
    __syncwarp();
  }
  if (cpos != nullptr) {
    for (int c = lane; c < m; c += 32) {
      const int j = sidx[c];
      float* d = cpos + ((size_t)o * m + c) * 3;
      d[0] = px[j];
      d[1] = py[j];
      d[2] = pz[j];
    }
  }
  if (!do_ball) return;

  const unsigned lt_mask = (1u << lane) - 1u;
  for (int c = 0; c < m; ++c) {
    const int jc = sidx[c];
    const float cx = px[jc], cy = py[jc], cz = pz[jc];
    int32_t* out = nbr + ((size_t)o * m + c) * T2P_MAX_NEIGHBORS;
    int count = 0;
    for (int j0 = 0; j0 < P && count < T2P_MAX_NEIGHBORS; j0 += 32) {
      const int j = j0 + lane;
      const bool in = (j < P) && (sqdist3_nofma(px[j < P ? j : 0], py[j < P ? j : 0], pz[j < P ? j : 0], cx, cy, cz) < r2);
      const unsigned b = __ballot_sync(0xffffffffu, in);
      const int slot = count + __popc(b & lt_mask);
      if (in && slot < T2P_MAX_NEIGHBORS) out[slot] = j;
      count += __popc(b);
    }
    count = min(count, T2P_MAX_NEIGHBORS);
    if (lane >= count) out[lane] = -1;
    if (lane == 0) cnt[(size_t)o * m + c] = count;
  }
}

int launch_fps_ball_mode(const float* pos, int n_obj, int P, int m, float r2, int mode, int do_ball, int32_t* ctr_idx,
                         float* cpos, int32_t* nbr, int32_t* cnt, cudaStream_t s) {
  if (n_obj <= 0) return T2P_OK;
  T2P_REQUIRE(P >= 1 && P <= 1024, T2P_ERR_UNSUPPORTED, "points per object P=%d outside [1,1024]", P);
  T2P_REQUIRE(m >= 1 && m <= P, T2P_ERR_INVALID, "fps: m=%d must be in [1,P=%d]", m, P);
  const size_t smem = (size_t)FB_WARPS * (3 * P + m) * sizeof(float);
  const int grid = (n_obj + FB_WARPS - 1) / FB_WARPS;
  const int ppl = (P + 31) / 32;
#define T2P_FB_LAUNCH(PPL_)                                                                              \
  do {                                                                                                   \
    if (smem > 48 * 1024)                                                                                \
      T2P_CUDA(cudaFuncSetAttribute(fps_ball_kernel<PPL_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    fps_ball_kernel<PPL_><<<grid, FB_WARPS * 32, smem, s>>>(pos, n_obj, P, m, r2, mode, do_ball, ctr_idx, cpos, nbr, cnt); \
  } while (0)
  if (ppl <= 1) T2P_FB_LAUNCH(1);
  else if (ppl <= 2) T2P_FB_LAUNCH(2);
  else if (ppl <= 4) T2P_FB_LAUNCH(4);
  else if (ppl <= 8) T2P_FB_LAUNCH(8);
  else if (ppl <= 16) T2P_FB_LAUNCH(16);
  else T2P_FB_LAUNCH(32);
#undef T2P_FB_LAUNCH
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

int launch_fps_ball(const float* pos, int n_obj, int P, int m, float r2, int32_t* ctr_idx, float* cpos, int32_t* nbr,
                    int32_t* cnt, cudaStream_t s) {
  return launch_fps_ball_mode(pos, n_obj, P, m, r2, 0, 1, ctr_idx, cpos, nbr, cnt, s);
}

// ---------------------------------------------------------------------------------------------------
// DynamicEdgeConv kNN: one CTA per cell, one thread per query object, embeddings in shared memory.
// Squared distance accumulated sequentially over the channels with individually rounded ops; the 8 best
// are kept by insertion with strict '<' so that ties keep the lower index (torch_cluster knn order).
// ---------------------------------------------------------------------------------------------------
constexpr int KNN_MAX_OBJ = 128;

__global__ void __launch_bounds__(KNN_MAX_OBJ)
knn_cells_kernel(const float* __restrict__ e, const int32_t* __restrict__ cell_offsets, int D, int cap,
                 int32_t* __restrict__ knn, int32_t* __restrict__ obj_cell) {
  extern __shared__ float ks[];  // [cap][D+1]
  const int cell = blockIdx.x;
  const int first = cell_offsets[cell];
  const int n = min(cell_offsets[cell + 1] - first, cap);  // host guarantees n <= cap (max_cell_objects)
  const int pitch = D + 1;
  for (int t = threadIdx.x; t < n * D; t += blockDim.x) {
    const int i = t / D, c = t - i * D;
    ks[i * pitch + c] = __ldg(e + (size_t)(first + i) * D + c);
  }
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= n) return;
  float bd[T2P_KNN_K];
  int bi[T2P_KNN_K];
#pragma unroll
  for (int s = 0; s < T2P_KNN_K; ++s) {
    bd[s] = __int_as_float(0x7f800000);
    bi[s] = -1;
  }
  const float* ei = ks + i * pitch;
  for (int j = 0; j < n; ++j) {
    const float* ej = ks + j * pitch;
    float acc = 0.f;
    for (int c = 0; c < D; ++c) {
      const float d = __fsub_rn(ei[c], ej[c]);
      acc = __fadd_rn(acc, __fmul_rn(d, d));
    }
    if (acc < bd[T2P_KNN_K - 1]) {
      bd[T2P_KNN_K - 1] = acc;
      bi[T2P_KNN_K - 1] = j;
#pragma unroll
      for (int p = T2P_KNN_K - 1; p > 0; --p) {
        if (bd[p] < bd[p - 1]) {
          const float td = bd[p]; bd[p] = bd[p - 1]; bd[p - 1] = td;
          const int ti = bi[p]; bi[p] = bi[p - 1]; bi[p - 1] = ti;
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < T2P_KNN_K; ++s) knn[(size_t)(first + i) * T2P_KNN_K + s] = bi[s] < 0 ? -1 : first + bi[s];
  obj_cell[first + i] = cell;
}

int launch_knn_cells(const float* e, const int32_t* cell_offsets, int n_cells, int max_cell_objects, int D,
                     int32_t* knn, int32_t* obj_cell, cudaStream_t s) {
  if (n_cells <= 0) return T2P_OK;
  T2P_REQUIRE(max_cell_objects >= 1 && max_cell_objects <= KNN_MAX_OBJ, T2P_ERR_UNSUPPORTED,
              "knn: max_cell_objects=%d outside [1,%d]", max_cell_objects, KNN_MAX_OBJ);
  const size_t smem = (size_t)max_cell_objects * (D + 1) * sizeof(float);
  T2P_REQUIRE(smem <= 227 * 1024, T2P_ERR_UNSUPPORTED, "knn: %d objects x embed_dim=%d exceed shared memory",
              max_cell_objects, D);
  if (smem > 48 * 1024)
    T2P_CUDA(cudaFuncSetAttribute(knn_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = (max_cell_objects + 31) / 32 * 32;
  knn_cells_kernel<<<n_cells, threads, smem, s>>>(e, cell_offsets, D, max_cell_objects, knn, obj_cell);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

}  // namespace t2p
