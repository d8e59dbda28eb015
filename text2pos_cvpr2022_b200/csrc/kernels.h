// Internal host-side launchers shared between translation units (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace t2p {

// y[M,N] (ld ldy) = act(x[M,K] (ld ldx) . W[K,N] + bias)
int launch_linear(const float* x, int M, int K, int ldx, const float* W, const float* bias, int N, bool relu,
                  float* y, int ldy, cudaStream_t s, const int32_t* run_if = nullptr);
// y = act([xa | xb] . W + bias): xa [M,Ka] (ld lda), xb [M,Kb] (ld ldb), K = Ka + Kb
int launch_linear_concat(const float* xa, int Ka, int lda, const float* xb, int Kb, int ldb, int M, const float* W,
                         const float* bias, int N, bool relu, float* y, int ldy, cudaStream_t s, const int32_t* run_if = nullptr);
// out[row / rows_per_group, N] = max over the group's rows of relu([xa|xb] . W + bias); out must be zeroed
int launch_linear_groupmax(const float* xa, int Ka, int lda, const float* xb, int Kb, int ldb, int M, const float* W,
                           const float* bias, int N, int rows_per_group, float* out, int ldo, cudaStream_t s,
                           const int32_t* run_if = nullptr);
// p[0..n) = 0 if *flag != 0 (device-side conditional, used with the `run_if` re-runs below)
int launch_zero_if(float* p, size_t n, const int32_t* flag, cudaStream_t s);
int launch_l2_normalize_rows(float* x, int M, int width, int ld, cudaStream_t s);

// PointConv message + second local_nn layer + max aggregation for one set-abstraction level:
//   out[o*m + c, :] = max over edges (j -> c) of relu(relu(T[j] - S[c]) . W2 + b2)
// T [n_obj*P, C1] (first layer applied to [x_j | pos_j], bias included), S [n_obj*m, C1] (= pos_c . W1p).
// Edges of centre c: its ball-query list (nbr/cnt) plus, with the quirk, flat point (lo*m + c) of the cell.
int launch_sa_edge(const float* T, const float* S, const int32_t* nbr, const int32_t* cnt,
                   const int32_t* obj_cell_start, int quirk, int n_obj, int P, int m, int C1, const float* W2,
                   const float* b2, int C2, float* out, cudaStream_t s, const int32_t* run_if = nullptr);
// `run_if` (device int, may be NULL): the kernel returns immediately unless *run_if != 0 -- the exact-fp32 re-run of a
// layer whose tensor-core pass flagged an activation outside the fp16 range.

// DynamicEdgeConv second layer + max over neighbours + max over the cell's objects:
//   pooled[cell(i), :] = max_i max_{j in knn(i)} relu(relu(AB[i, :D] + AB[j, D:]) . W2 + b2)
int launch_edgeconv(const float* AB, const int32_t* knn, const int32_t* obj_cell, int n_obj, int D, const float* W2,
                    const float* b2, float* pooled, cudaStream_t s);

// set abstraction second layer on the tensor cores (C1 = C2 in {128, 256}), csrc/sa_tc.cu
bool sa_edge_tc_supported(int C1, int C2, int m);
int launch_sa_edge_tc(const float* T, const float* S, const int32_t* nbr, const int32_t* cnt, const int32_t* obj_cell_start,
                      int quirk, int n_obj, int P, int m, int C, const float* w_img, const float* b2, float* out, int sms,
                      int32_t* overflow_flag, cudaStream_t s);
bool linear_groupmax_tc_supported(int K, int N);
// y[M, N] = act(x[M, K] (ld ldx, x >= 0) . W + bias + pos[M, 3] . Wp) on the tensor cores (w_img: fp16 hi/lo images of 2^8 W per
// column block, Wp: fp32 [3, N] or NULL); supported (K, N): the dense layers of PointNet++ (see linear_tc_supported)
bool linear_tc_supported(int K, int N);
int launch_linear_tc(const float* x, int ldx, int M, int K, const float* w_img, const float* bias, int N, const float* pos,
                     const float* Wp, bool relu, float* y, int sms, int32_t* overflow_flag, cudaStream_t s);
int launch_linear_groupmax_tc(const float* x, int M, int K, const float* w_img, const float* bias, int N, int group, float* out,
                              int sms, int32_t* overflow_flag, cudaStream_t s);
int launch_fps_ball(const float* pos, int n_obj, int P, int m, float r2, int32_t* ctr_idx, float* cpos,
                    int32_t* nbr, int32_t* cnt, cudaStream_t s);
int launch_fps_ball_mode(const float* pos, int n_obj, int P, int m, float r2, int mode, int do_ball, int32_t* ctr_idx,
                         float* cpos, int32_t* nbr, int32_t* cnt, cudaStream_t s);
int launch_knn_cells(const float* e, const int32_t* cell_offsets, int n_cells, int max_cell_objects, int D,
                     int32_t* knn, int32_t* obj_cell, cudaStream_t s);


// tensor-core (tcgen05 + TMA) retrieval path, csrc/retrieval_tc.cu
struct TcPlan {
  bool ok;
  int KP, NC, tile_n, tiles, G, tiles_per_cta, nb_bits, qtiles, dup, qstream, nsrc, q_pitch, st_pitch, stages, a_tmem, slack, parts;
  size_t scan_smem, sel_smem;
};
TcPlan tc_plan(int B, int N, int D, int k, int sms);
size_t tc_workspace_bytes(const TcPlan& p, int B);
int launch_row_norm2_max(const float* db, int N, int D, float* out, cudaStream_t s);
int launch_retrieve_tc(const TcPlan& p, const float* d_q, const float* d_db, int B, int N, int D, int k, int64_t idx_base,
                       const float* d_db_norm2_max, int force_rescan, double* d_out_scores, int64_t* d_out_idx, int32_t* d_stats,
                       void* d_ws, size_t ws_bytes, cudaStream_t s);


// SuperGlue head on the tensor cores (D = 128), csrc/superglue_tc.cu
bool superglue_tc_supported(const t2p_superglue_desc* desc, int M, int N);
int launch_superglue_tc(const float* blob, const t2p_superglue_desc* desc, const float* d_desc0, const int64_t* d_idx0,
                        const float* d_desc1, const int64_t* d_idx1, int B, int M, int N, float* d_P, int64_t* d_matches0,
                        int64_t* d_matches1, float* d_mscores0, float* d_mscores1, float* d_dbg_scores, int32_t* overflow_flag,
                        cudaStream_t s);

// tensor-core LSTM recurrence (H = 256), csrc/lstm_tc.cu
int launch_lstm_tc(const float* xproj4, const float* w_img, const int32_t* tokens, const int32_t* lengths, int B, int T, int V,
                   float* hfinal, int max_groups, cudaStream_t s);

}  // namespace t2p
