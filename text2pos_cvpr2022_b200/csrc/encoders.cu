// Host-side orchestration of the cell encoder (PointNet++ -> object embedding -> cell aggregation) and the
// C-ABI wrappers of the primitives.  Everything is enqueued on the caller's stream; no synchronisation.
#include "kernels.h"

using namespace t2p;

namespace {

struct PnDims {
  int Pd[4];  // points per object at the input of level l (Pd[3] = survivors)
  int Cin[3], C1[3], C2[3];
  int ga_h, ga_o, f1, f2;
};

static int pn_dims(const t2p_pointnet2_desc* d, int P, PnDims* o) {
  o->Pd[0] = P;
  for (int l = 0; l < 3; ++l) {
    o->Pd[l + 1] = (o->Pd[l] + 1) / 2;  // ceil(0.5 * n), torch_cluster fps
    o->Cin[l] = d->sa_l1[l].k - 3;
    o->C1[l] = d->sa_l1[l].n;
    o->C2[l] = d->sa_l2[l].n;
    T2P_REQUIRE(o->Cin[l] >= 1 && d->sa_l2[l].k == o->C1[l], T2P_ERR_INVALID, "pointnet2: inconsistent sa%d dims", l + 1);
    if (l > 0) T2P_REQUIRE(o->Cin[l] == o->C2[l - 1], T2P_ERR_INVALID, "pointnet2: sa%d input != sa%d output", l + 1, l);
  }
  T2P_REQUIRE(o->Cin[0] == 3, T2P_ERR_INVALID, "pointnet2: sa1 expects rgb (3) + xyz (3) inputs");
  T2P_REQUIRE(d->ga_l1.k == o->C2[2] + 3 && d->ga_l2.k == d->ga_l1.n && d->lin1.k == d->ga_l2.n && d->lin2.k == d->lin1.n,
              T2P_ERR_INVALID, "pointnet2: inconsistent ga/lin dims");
  o->ga_h = d->ga_l1.n;
  o->ga_o = d->ga_l2.n;
  o->f1 = d->lin1.n;
  o->f2 = d->lin2.n;
  return T2P_OK;
}

struct PnWorkspace {
  int32_t *ctr_idx, *nbr, *cnt;
  float *cpos[3], *T, *S, *x[3], *gah, *f0, *f1;
  int32_t* flags;  // [8] fp16 range flags of the tensor-core layers (SA1..3, global abstraction)
};

static size_t pn_carve(const PnDims& d, int n_obj, Arena& a, PnWorkspace* w) {
  const size_t n = (size_t)n_obj;
  size_t tmax = 0, smax = 0;
  for (int l = 0; l < 3; ++l) {
    tmax = std::max(tmax, (size_t)d.Pd[l] * d.C1[l]);
    smax = std::max(smax, (size_t)d.Pd[l + 1] * d.C1[l]);
  }
  w->ctr_idx = a.take<int32_t>(n * d.Pd[1]);
  w->nbr = a.take<int32_t>(n * d.Pd[1] * T2P_MAX_NEIGHBORS);
  w->cnt = a.take<int32_t>(n * d.Pd[1]);
  for (int l = 0; l < 3; ++l) w->cpos[l] = a.take<float>(n * d.Pd[l + 1] * 3);
  w->T = a.take<float>(n * tmax);
  w->S = a.take<float>(n * smax);
  for (int l = 0; l < 3; ++l) w->x[l] = a.take<float>(n * d.Pd[l + 1] * d.C2[l]);
  w->gah = a.take<float>(n * d.Pd[3] * d.ga_h);
  w->f0 = a.take<float>(n * d.ga_o);
  w->f1 = a.take<float>(n * d.f1);
  w->flags = a.take<int32_t>(16);
  return a.used;
}

}  // namespace

namespace {
int sm_count_cached() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess) sms = prop.multiProcessorCount;
    if (sms <= 0) sms = 148;
  }
  return sms;
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------------
// primitives
// ---------------------------------------------------------------------------------------------------
int t2p_fps(const float* d_pos, int n_obj, int P, int m, int32_t* d_idx, t2p_stream stream) {
  T2P_REQUIRE(d_pos && d_idx && n_obj >= 0, T2P_ERR_INVALID, "fps: null argument");
  return launch_fps_ball_mode(d_pos, n_obj, P, m, 0.f, 0, 0, d_idx, nullptr, nullptr, nullptr, as_stream(stream));
}

int t2p_ball_query(const float* d_pos, const int32_t* d_ctr_idx, int n_obj, int P, int m, float r2, int cap,
                   int32_t* d_nbr, int32_t* d_count, t2p_stream stream) {
  T2P_REQUIRE(d_pos && d_ctr_idx && d_nbr && d_count && n_obj >= 0, T2P_ERR_INVALID, "ball_query: null argument");
  T2P_REQUIRE(cap == T2P_MAX_NEIGHBORS, T2P_ERR_UNSUPPORTED, "ball_query: cap must be %d", T2P_MAX_NEIGHBORS);
  return launch_fps_ball_mode(d_pos, n_obj, P, m, r2, 1, 1, const_cast<int32_t*>(d_ctr_idx), nullptr, d_nbr, d_count,
                              as_stream(stream));
}

int t2p_linear(const t2p_weights* w, const t2p_linear_desc* lin, const float* d_x, int M, int ldx, int relu, float* d_y,
               int ldy, t2p_stream stream) {
  T2P_REQUIRE(w && lin && d_x && d_y, T2P_ERR_INVALID, "linear: null argument");
  T2P_REQUIRE(lin_ok(w, *lin), T2P_ERR_INVALID, "linear: descriptor outside the weight blob");
  T2P_REQUIRE(ldx >= lin->k && ldy >= lin->n, T2P_ERR_INVALID, "linear: leading dimension too small");
  return launch_linear(d_x, M, lin->k, ldx, wptr(w, lin->w_off), wptr(w, lin->b_off), lin->n, relu != 0, d_y, ldy,
                       as_stream(stream));
}

int t2p_l2_normalize_rows(float* d_x, int M, int width, int ld, t2p_stream stream) {
  T2P_REQUIRE(d_x && width > 0 && ld >= width, T2P_ERR_INVALID, "l2_normalize_rows: bad argument");
  return launch_l2_normalize_rows(d_x, M, width, ld, as_stream(stream));
}

int t2p_knn_cells(const float* d_e, const int32_t* d_cell_offsets, int n_obj, int n_cells, int max_cell_objects, int D,
                  int32_t* d_knn, int32_t* d_obj_cell, t2p_stream stream) {
  T2P_REQUIRE(d_e && d_cell_offsets && d_knn && d_obj_cell && n_obj >= 0, T2P_ERR_INVALID, "knn_cells: null argument");
  return launch_knn_cells(d_e, d_cell_offsets, n_cells, max_cell_objects, D, d_knn, d_obj_cell, as_stream(stream));
}

// ---------------------------------------------------------------------------------------------------
// PointNet++
// ---------------------------------------------------------------------------------------------------
size_t t2p_pointnet2_workspace(const t2p_pointnet2_desc* desc, int n_obj, int P) {
  PnDims d;
  if (!desc || n_obj <= 0 || P <= 0 || pn_dims(desc, P, &d) != T2P_OK) return 0;
  Arena a(nullptr, 0);
  PnWorkspace w;
  return pn_carve(d, n_obj, a, &w);
}

int t2p_pointnet2_forward(const t2p_weights* w, const t2p_pointnet2_desc* desc, const float* d_pos, const float* d_rgb,
                          const int32_t* d_obj_cell_start, int n_obj, int P, float* d_features2,
                          int32_t* const* d_dbg_idx, int32_t* const* d_dbg_nbr, int32_t* const* d_dbg_cnt,
                          float* const* d_dbg_x, void* d_ws, size_t ws_bytes, t2p_stream stream) {
  T2P_REQUIRE(w && desc && d_pos && d_rgb && d_features2, T2P_ERR_INVALID, "pointnet2: null argument");
  T2P_REQUIRE(!desc->self_loop_quirk || d_obj_cell_start, T2P_ERR_INVALID, "pointnet2: self_loop_quirk needs d_obj_cell_start");
  if (n_obj <= 0) return T2P_OK;
  PnDims d;
  T2P_TRY(pn_dims(desc, P, &d));
  for (int l = 0; l < 3; ++l)
    T2P_REQUIRE(lin_ok(w, desc->sa_l1[l]) && lin_ok(w, desc->sa_l2[l]), T2P_ERR_INVALID, "pointnet2: sa%d outside blob", l + 1);
  T2P_REQUIRE(lin_ok(w, desc->ga_l1) && lin_ok(w, desc->ga_l2) && lin_ok(w, desc->lin1) && lin_ok(w, desc->lin2),
              T2P_ERR_INVALID, "pointnet2: ga/lin outside blob");
  Arena a(d_ws, ws_bytes);
  PnWorkspace ws;
  pn_carve(d, n_obj, a, &ws);
  T2P_REQUIRE(a.ok, T2P_ERR_WORKSPACE, "pointnet2: workspace %zu < %zu bytes", ws_bytes, a.used);
  cudaStream_t s = as_stream(stream);
  T2P_CUDA(cudaMemsetAsync(ws.flags, 0, 16 * sizeof(int32_t), s));  // fp16 range flags of the tensor-core layers
  const int sms = sm_count_cached();
  // dense layer on the tensor cores when its images are packed (desc->dense_tc_off[i] >= 0), with the exact-fp32 kernel behind it as
  // the conditional re-run for activations outside the fp16 range; else the fp32 kernel alone
  auto dense = [&](int slot, const float* x, int K, int M, const t2p_linear_desc& lin, const float* pos, bool relu, float* y) -> int {
    const float* W = wptr(w, lin.w_off);
    const float* bias = wptr(w, lin.b_off);
    const int N = lin.n;
    const int32_t* run_if = nullptr;
    if (desc->dense_tc_off[slot] >= 0 && linear_tc_supported(K, N) && (size_t)desc->dense_tc_off[slot] + (size_t)K * N <= w->n_floats) {
      T2P_TRY(launch_linear_tc(x, K, M, K, wptr(w, desc->dense_tc_off[slot]), bias, N, pos, pos ? W + (size_t)K * N : nullptr, relu, y,
                               sms, ws.flags + 4 + slot, s));
      run_if = ws.flags + 4 + slot;
    }
    if (pos) return launch_linear_concat(x, K, K, pos, 3, 3, M, W, bias, N, relu, y, N, s, run_if);
    return launch_linear(x, M, K, K, W, bias, N, relu, y, N, s, run_if);
  };

  const float* x_in = d_rgb;
  const float* pos_in = d_pos;
  for (int l = 0; l < 3; ++l) {
    const int Pd = d.Pd[l], m = d.Pd[l + 1], Cin = d.Cin[l], C1 = d.C1[l], C2 = d.C2[l];
    const float* W1 = wptr(w, desc->sa_l1[l].w_off);
    T2P_TRY(launch_fps_ball(pos_in, n_obj, Pd, m, desc->sa_radius_sq[l], ws.ctr_idx, ws.cpos[l], ws.nbr, ws.cnt, s));
    // T_j = W1 . [x_j | pos_j] + b1 for every point;  S_c = W1[pos rows] . pos_c for every centre
    if (l >= 1) {
      T2P_TRY(dense(l - 1, x_in, Cin, n_obj * Pd, desc->sa_l1[l], pos_in, false, ws.T));
    } else {
      T2P_TRY(launch_linear_concat(x_in, Cin, Cin, pos_in, 3, 3, n_obj * Pd, W1, wptr(w, desc->sa_l1[l].b_off), C1, false,
                                   ws.T, C1, s));
    }
    T2P_TRY(launch_linear(ws.cpos[l], n_obj * m, 3, 3, W1 + (size_t)Cin * C1, nullptr, C1, false, ws.S, C1, s));
    T2P_CUDA(cudaMemsetAsync(ws.x[l], 0, (size_t)n_obj * m * C2 * sizeof(float), s));
    if (desc->sa_l2_tc_off[l] >= 0 && sa_edge_tc_supported(C1, C2, m) &&
        (size_t)desc->sa_l2_tc_off[l] + (size_t)((C1 + 63) / 64 * 64) * ((C2 + 127) / 128 * 128) <= w->n_floats) {
      // second layer + ReLU + max on the tensor cores (fp16 hi/lo split, fp32 accumulate).  If an activation did not fit the
      // fp16 range the kernel raises ws.flags[l] and the two launches behind it redo the layer in exact fp32 (they return
      // at once otherwise): results stay within the 1e-4 contract for any weights, at tensor-core speed for sane ones.
      T2P_TRY(launch_sa_edge_tc(ws.T, ws.S, ws.nbr, ws.cnt, d_obj_cell_start, desc->self_loop_quirk, n_obj, Pd, m, C1,
                                wptr(w, desc->sa_l2_tc_off[l]), wptr(w, desc->sa_l2[l].b_off), ws.x[l], sm_count_cached(),
                                ws.flags + l, s));
      T2P_TRY(launch_zero_if(ws.x[l], (size_t)n_obj * m * C2, ws.flags + l, s));
      T2P_TRY(launch_sa_edge(ws.T, ws.S, ws.nbr, ws.cnt, d_obj_cell_start, desc->self_loop_quirk, n_obj, Pd, m, C1,
                             wptr(w, desc->sa_l2[l].w_off), wptr(w, desc->sa_l2[l].b_off), C2, ws.x[l], s, ws.flags + l));
    } else {
      T2P_TRY(launch_sa_edge(ws.T, ws.S, ws.nbr, ws.cnt, d_obj_cell_start, desc->self_loop_quirk, n_obj, Pd, m, C1,
                             wptr(w, desc->sa_l2[l].w_off), wptr(w, desc->sa_l2[l].b_off), C2, ws.x[l], s));
    }
    if (d_dbg_idx && d_dbg_idx[l])
      T2P_CUDA(cudaMemcpyAsync(d_dbg_idx[l], ws.ctr_idx, (size_t)n_obj * m * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    if (d_dbg_nbr && d_dbg_nbr[l])
      T2P_CUDA(cudaMemcpyAsync(d_dbg_nbr[l], ws.nbr, (size_t)n_obj * m * T2P_MAX_NEIGHBORS * sizeof(int32_t),
                               cudaMemcpyDeviceToDevice, s));
    if (d_dbg_cnt && d_dbg_cnt[l])
      T2P_CUDA(cudaMemcpyAsync(d_dbg_cnt[l], ws.cnt, (size_t)n_obj * m * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    if (d_dbg_x && d_dbg_x[l])
      T2P_CUDA(cudaMemcpyAsync(d_dbg_x[l], ws.x[l], (size_t)n_obj * m * C2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    x_in = ws.x[l];
    pos_in = ws.cpos[l];
  }
  // global abstraction: get_mlp([259,512,1024]) on cat(x,pos) then per-object max; lin1, lin2 with ReLU
  const int m3 = d.Pd[3];
  T2P_TRY(dense(2, x_in, d.C2[2], n_obj * m3, desc->ga_l1, pos_in, true, ws.gah));
  T2P_CUDA(cudaMemsetAsync(ws.f0, 0, (size_t)n_obj * d.ga_o * sizeof(float), s));
  if (desc->ga_l2_tc_off >= 0 && linear_groupmax_tc_supported(d.ga_h, d.ga_o) &&
      (size_t)desc->ga_l2_tc_off + (size_t)d.ga_h * d.ga_o <= w->n_floats) {
    T2P_TRY(launch_linear_groupmax_tc(ws.gah, n_obj * m3, d.ga_h, wptr(w, desc->ga_l2_tc_off), wptr(w, desc->ga_l2.b_off), d.ga_o,
                                      m3, ws.f0, sm_count_cached(), ws.flags + 3, s));
    T2P_TRY(launch_zero_if(ws.f0, (size_t)n_obj * d.ga_o, ws.flags + 3, s));
    T2P_TRY(launch_linear_groupmax(ws.gah, d.ga_h, d.ga_h, nullptr, 0, 0, n_obj * m3, wptr(w, desc->ga_l2.w_off),
                                   wptr(w, desc->ga_l2.b_off), d.ga_o, m3, ws.f0, d.ga_o, s, ws.flags + 3));
  } else {
    T2P_TRY(launch_linear_groupmax(ws.gah, d.ga_h, d.ga_h, nullptr, 0, 0, n_obj * m3, wptr(w, desc->ga_l2.w_off),
                                   wptr(w, desc->ga_l2.b_off), d.ga_o, m3, ws.f0, d.ga_o, s));
  }
  T2P_TRY(dense(3, ws.f0, d.ga_o, n_obj, desc->lin1, nullptr, true, ws.f1));
  T2P_TRY(dense(4, ws.f1, d.f1, n_obj, desc->lin2, nullptr, true, d_features2));
  return T2P_OK;
}

// ---------------------------------------------------------------------------------------------------
// object embedding
// ---------------------------------------------------------------------------------------------------
size_t t2p_object_embed_workspace(const t2p_objenc_desc* desc, int n_obj) {
  if (!desc || n_obj <= 0) return 0;
  Arena a(nullptr, 0);
  a.take<float>((size_t)n_obj * 3 * desc->embed_dim);
  a.take<float>((size_t)n_obj * std::max(desc->color_l1.n, desc->pos_l1.n));
  return a.used;
}

int t2p_object_embed(const t2p_weights* w, const t2p_objenc_desc* desc, const float* d_features2, const float* d_centers,
                     const float* d_mean_rgb, int n_obj, float* d_emb, void* d_ws, size_t ws_bytes, t2p_stream stream) {
  T2P_REQUIRE(w && desc && d_features2 && d_centers && d_mean_rgb && d_emb, T2P_ERR_INVALID, "object_embed: null argument");
  if (n_obj <= 0) return T2P_OK;
  const int D = desc->embed_dim;
  const t2p_linear_desc* all[] = {&desc->mlp_pointnet, &desc->color_l1, &desc->color_l2, &desc->pos_l1, &desc->pos_l2, &desc->merge};
  for (const t2p_linear_desc* l : all) T2P_REQUIRE(lin_ok(w, *l), T2P_ERR_INVALID, "object_embed: descriptor outside blob");
  T2P_REQUIRE(desc->mlp_pointnet.n == D && desc->color_l2.n == D && desc->pos_l2.n == D && desc->merge.k == 3 * D &&
                  desc->merge.n == D && desc->color_l1.k == 3 && desc->pos_l1.k == 3 &&
                  desc->color_l2.k == desc->color_l1.n && desc->pos_l2.k == desc->pos_l1.n,
              T2P_ERR_INVALID, "object_embed: inconsistent dims");
  Arena a(d_ws, ws_bytes);
  float* cat = a.take<float>((size_t)n_obj * 3 * D);
  float* hid = a.take<float>((size_t)n_obj * std::max(desc->color_l1.n, desc->pos_l1.n));
  T2P_REQUIRE(a.ok, T2P_ERR_WORKSPACE, "object_embed: workspace %zu < %zu bytes", ws_bytes, a.used);
  cudaStream_t s = as_stream(stream);
  auto lin = [&](const t2p_linear_desc& l, const float* x, int ldx, float* y, int ldy) {
    return launch_linear(x, n_obj, l.k, ldx, wptr(w, l.w_off), wptr(w, l.b_off), l.n, true, y, ldy, s);
  };
  // class slot: normalize(mlp_pointnet(features2))        object_encoder.py:98,111
  T2P_TRY(lin(desc->mlp_pointnet, d_features2, desc->mlp_pointnet.k, cat, 3 * D));
  T2P_TRY(launch_l2_normalize_rows(cat, n_obj, D, 3 * D, s));
  // colour slot: normalize(color_encoder(mean rgb))         :121-127
  T2P_TRY(lin(desc->color_l1, d_mean_rgb, 3, hid, desc->color_l1.n));
  T2P_TRY(lin(desc->color_l2, hid, desc->color_l1.n, cat + D, 3 * D));
  T2P_TRY(launch_l2_normalize_rows(cat + D, n_obj, D, 3 * D, s));
  // position slot: normalize(pos_encoder(centre))           :129-135
  T2P_TRY(lin(desc->pos_l1, d_centers, 3, hid, desc->pos_l1.n));
  T2P_TRY(lin(desc->pos_l2, hid, desc->pos_l1.n, cat + 2 * D, 3 * D));
  T2P_TRY(launch_l2_normalize_rows(cat + 2 * D, n_obj, D, 3 * D, s));
  // merge                                                    :138
  T2P_TRY(lin(desc->merge, cat, 3 * D, d_emb, D));
  return T2P_OK;
}

// ---------------------------------------------------------------------------------------------------
// cell aggregation
// ---------------------------------------------------------------------------------------------------
namespace {
struct CaWorkspace {
  float *e, *AB, *pooled, *h;
  int32_t *knn, *obj_cell;
};
static size_t ca_carve(int D, int n_obj, int n_cells, Arena& a, CaWorkspace* w) {
  w->e = a.take<float>((size_t)n_obj * D);
  w->AB = a.take<float>((size_t)n_obj * 2 * D);
  w->pooled = a.take<float>((size_t)n_cells * D);
  w->h = a.take<float>((size_t)n_cells * D);
  w->knn = a.take<int32_t>((size_t)n_obj * T2P_KNN_K);
  w->obj_cell = a.take<int32_t>((size_t)n_obj);
  return a.used;
}
}  // namespace

size_t t2p_cell_aggregate_workspace(const t2p_cellagg_desc* desc, int n_obj, int n_cells) {
  if (!desc || n_obj <= 0 || n_cells <= 0) return 0;
  Arena a(nullptr, 0);
  CaWorkspace w;
  return ca_carve(desc->embed_dim, n_obj, n_cells, a, &w);
}

int t2p_cell_aggregate(const t2p_weights* w, const t2p_cellagg_desc* desc, const float* d_emb,
                       const int32_t* d_cell_offsets, int n_obj, int n_cells, int max_cell_objects, float* d_out,
                       int32_t* d_dbg_knn, void* d_ws, size_t ws_bytes, t2p_stream stream) {
  T2P_REQUIRE(w && desc && d_emb && d_cell_offsets && d_out, T2P_ERR_INVALID, "cell_aggregate: null argument");
  if (n_obj <= 0 || n_cells <= 0) return T2P_OK;
  const int D = desc->embed_dim;
  T2P_REQUIRE(lin_ok(w, desc->edge_ab) && lin_ok(w, desc->edge_l2) && lin_ok(w, desc->lin_l1) && lin_ok(w, desc->lin_l2),
              T2P_ERR_INVALID, "cell_aggregate: descriptor outside blob");
  T2P_REQUIRE(desc->edge_ab.k == D && desc->edge_ab.n == 2 * D && desc->edge_l2.k == D && desc->edge_l2.n == D &&
                  desc->lin_l1.k == D && desc->lin_l2.k == desc->lin_l1.n && desc->lin_l2.n == D && desc->lin_l1.n == D,
              T2P_ERR_INVALID, "cell_aggregate: inconsistent dims");
  Arena a(d_ws, ws_bytes);
  CaWorkspace ws;
  ca_carve(D, n_obj, n_cells, a, &ws);
  T2P_REQUIRE(a.ok, T2P_ERR_WORKSPACE, "cell_aggregate: workspace %zu < %zu bytes", ws_bytes, a.used);
  cudaStream_t s = as_stream(stream);
  // e = F.normalize(emb)                                   cell_retrieval.py:94
  T2P_CUDA(cudaMemcpyAsync(ws.e, d_emb, (size_t)n_obj * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
  T2P_TRY(launch_l2_normalize_rows(ws.e, n_obj, D, D, s));
  // DynamicEdgeConv(k=8, max) + global_max_pool              :97-98
  T2P_TRY(launch_knn_cells(ws.e, d_cell_offsets, n_cells, max_cell_objects, D, ws.knn, ws.obj_cell, s));
  if (d_dbg_knn)
    T2P_CUDA(cudaMemcpyAsync(d_dbg_knn, ws.knn, (size_t)n_obj * T2P_KNN_K * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  T2P_TRY(launch_linear(ws.e, n_obj, D, D, wptr(w, desc->edge_ab.w_off), wptr(w, desc->edge_ab.b_off), 2 * D, false,
                        ws.AB, 2 * D, s));
  T2P_CUDA(cudaMemsetAsync(ws.pooled, 0, (size_t)n_cells * D * sizeof(float), s));
  T2P_TRY(launch_edgeconv(ws.AB, ws.knn, ws.obj_cell, n_obj, D, wptr(w, desc->edge_l2.w_off),
                          wptr(w, desc->edge_l2.b_off), ws.pooled, s));
  // lin (get_mlp with BN and trailing ReLU) + normalize       :99,105
  T2P_TRY(launch_linear(ws.pooled, n_cells, D, D, wptr(w, desc->lin_l1.w_off), wptr(w, desc->lin_l1.b_off), D, true, ws.h, D, s));
  T2P_TRY(launch_linear(ws.h, n_cells, D, D, wptr(w, desc->lin_l2.w_off), wptr(w, desc->lin_l2.b_off), D, true, d_out, D, s));
  T2P_TRY(launch_l2_normalize_rows(d_out, n_cells, D, D, s));
  return T2P_OK;
}

}  // extern "C"
