/*
 * text2pos_b200 -- C ABI of the B200-native Text2Pos hot path (sm_100a).
 *
 * The reference (mako443/Text2Pos-CVPR2022) is pure Python: it has no FFI layer.  Its only seam for
 * this path is the nn.Module API (models/cell_retrieval.py:69-107, models/superglue_matcher.py:87-128,
 * training/coarse.py:134-148).  Each entry point below replaces the library calls that one of those
 * Python call sites makes today; the Python modules in text2pos_cvpr2022_b200/ (and the `models.*`
 * shims) bind them with ctypes, see INTEGRATION.md.
 *
 * Conventions
 *  - every pointer named d_* is a DEVICE pointer owned by the caller; h_* is a HOST pointer;
 *  - all calls are asynchronous on `stream` (a cudaStream_t passed as void*), never synchronise the
 *    device, never free caller memory, hold no global mutable state (safe from several host threads on
 *    distinct streams);
 *  - return value: 0 = ok, <0 = error (T2P_ERR_*); t2p_last_error() gives a thread-local message;
 *  - scratch memory is caller-provided: ask t2p_*_workspace() for the size;
 *  - matrices are row-major, float32 unless stated; indices int32 unless stated.
 */
#ifndef TEXT2POS_B200_H
#define TEXT2POS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T2P_OK 0
#define T2P_ERR_INVALID (-1)     /* bad argument */
#define T2P_ERR_CUDA (-2)        /* CUDA runtime error */
#define T2P_ERR_WORKSPACE (-3)   /* workspace too small */
#define T2P_ERR_UNSUPPORTED (-4) /* shape outside the compiled range / not an sm_100 device */

#define T2P_MAX_NEIGHBORS 32 /* torch_cluster radius() default max_num_neighbors (pointnet2.py:28-30) */
#define T2P_KNN_K 8          /* DynamicEdgeConv k (cell_retrieval.py:46-48) */
#define T2P_MAX_GNN_LAYERS 32

typedef void* t2p_stream;

int t2p_version(void);
const char* t2p_last_error(void);
/* sm count / compute capability of the current device; T2P_ERR_UNSUPPORTED unless cc 10.x */
int t2p_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * Weights.  One immutable device copy of a packed float32 blob; the *_desc structs below address
 * sub-arrays of it by offset (in floats).  Packing (BatchNorm folding, transposition to [K,N]) is done
 * once by the host side from the reference's state_dict (text2pos_cvpr2022_b200/packing.py).
 * ------------------------------------------------------------------------------------------------ */
typedef struct t2p_weights t2p_weights;
int t2p_weights_create(const float* h_blob, size_t n_floats, t2p_weights** out);
int t2p_weights_destroy(t2p_weights* w);
const float* t2p_weights_device_ptr(const t2p_weights* w);

/* y = act(x . wT + bias): wT is [k, n] row-major at w_off, bias [n] at b_off (b_off < 0: no bias) */
typedef struct {
  int64_t w_off;
  int64_t b_off;
  int32_t k;
  int32_t n;
} t2p_linear_desc;

/* ------------------------------------------------------------------------------------------------
 * (a7) all-pairs scores + top-k.  Replaces training/coarse.py:134-148 (float64 numpy mat-vec + argsort).
 * d_q [B,D], d_db [N,D] float32.  Output ordered by (score desc, index asc); scores are the float64
 * dot products of the float32 inputs (the reference ranks in float64), indices are idx_base + row.
 * If N < k the tail is filled with index -1 / score -inf.
 * ------------------------------------------------------------------------------------------------ */
size_t t2p_retrieve_topk_workspace(int B, int N, int D, int k);
int t2p_retrieve_topk(const float* d_q, const float* d_db, int B, int N, int D, int k, int64_t idx_base,
                      double* d_out_scores, int64_t* d_out_idx, void* d_ws, size_t ws_bytes, t2p_stream stream);
/* Same, with the knobs of the tensor-core path.  When D % 32 == 0, D <= 256 and k <= 26 the scores are first
 * computed on the tcgen05 tensor cores in TF32 (TMA-fed scan with an in-kernel top-k); the candidates are then
 * re-scored in float64 and the result is CERTIFIED against the TF32 error bound eps*|q|*max|d|, with an exact
 * float64 rescan of the DB for any query that cannot be certified -- the output is always the float64 ranking.
 * Other shapes take the CUDA-core scan: float32 scores pre-select, float64 re-scores, and the same kind of certificate (float32
 * error bound) decides whether a query needs the float64 rescan.
 *   d_db_norm2_max: device scalar = max squared row norm of the DB (t2p_db_row_norm2_max, computed once per DB);
 *                   NULL = compute it inside the call (one extra pass over the DB);
 *   flags:          T2P_RETRIEVE_*;
 *   d_stats:        optional device int32[2]: [0] += queries certified from the fast scan, [1] += queries rescanned exactly. */
#define T2P_RETRIEVE_FORCE_GENERIC 1 /* always use the CUDA-core scan */
#define T2P_RETRIEVE_FORCE_RESCAN 2  /* tensor path, but treat every query as uncertified (tests the rescan) */
#define T2P_RETRIEVE_MAX_CTAS(n) (((n) & 0xff) << 8) /* tensor path: at most n scan CTAs in the whole launch, all query tiles
                                                        together (0 = one per SM).  Fewer CTAs = less SM-time per batch at more
                                                        latency: for servers with batches in flight */
int t2p_retrieve_topk_ex(const float* d_q, const float* d_db, int B, int N, int D, int k, int64_t idx_base,
                         const float* d_db_norm2_max, int flags, double* d_out_scores, int64_t* d_out_idx,
                         int32_t* d_stats, void* d_ws, size_t ws_bytes, t2p_stream stream);

int t2p_db_row_norm2_max(const float* d_db, int N, int D, float* d_out, t2p_stream stream);
/* merge R per-shard lists [R,B,k_in] (e.g. after an all-gather) into [B,k_out], same ordering rule */
int t2p_topk_merge(const double* d_scores, const int64_t* d_idx, int R, int B, int k_in, int k_out,
                   double* d_out_scores, int64_t* d_out_idx, t2p_stream stream);

/* ------------------------------------------------------------------------------------------------
 * Peer exchange of the sharded retrieval over NVLink peer memory, no NCCL on the data path.  Every rank
 * owns one buffer that its peers have mapped (CUDA IPC); `base[j]` is rank j's buffer as seen from THIS process.
 * t2p_peer_push: for every peer j, copies bytes0 from d_src0 + j*src_stride0 to base[j] + dst_off0 (and optionally a
 *   second block), then raises the 64-bit flag at base[j] + flag_off to (*d_epoch + 1).
 * t2p_peer_wait: spins (on the device) until the n_peers local flags at d_flags are all > *d_epoch, then increments
 *   *d_epoch.  Epochs live on the device and advance in step on every rank, so push/wait pairs can be captured in CUDA
 *   graphs.  A peer that never arrives traps after ~30 s (CUDA error at the caller) instead of hanging.
 * ------------------------------------------------------------------------------------------------ */
#define T2P_MAX_PEERS 8
typedef struct {
  void* base[T2P_MAX_PEERS];
  int32_t n_peers;
  int32_t my_rank;
} t2p_peers;
int t2p_enable_peer_access(int peer_device);
/* symmetric buffers: a zero-filled cudaMalloc block + its 64-byte CUDA IPC handle (the host side ships the handles to the
 * peers, e.g. with all_gather_object); t2p_ipc_open maps a peer's block into this process (call it with the CONSUMER's
 * device current: peer access is enabled lazily).  The owner frees with t2p_ipc_free after the peers have closed. */
int t2p_ipc_alloc(size_t bytes, void** d_ptr, void* handle64);
int t2p_ipc_open(const void* handle64, void** d_ptr);
int t2p_ipc_close(void* d_ptr);
int t2p_ipc_free(void* d_ptr);
int t2p_peer_push(const t2p_peers* peers, const void* d_src0, size_t src_stride0, size_t dst_off0, size_t bytes0,
                  const void* d_src1, size_t src_stride1, size_t dst_off1, size_t bytes1, size_t flag_off,
                  const uint64_t* d_epoch, t2p_stream stream);
int t2p_peer_wait(uint64_t* d_flags, int n_peers, uint64_t* d_epoch, t2p_stream stream);

/* ------------------------------------------------------------------------------------------------
 * (a2) PointNet++ primitives, exposed for bit-exact index parity.
 * Replaces torch_geometric fps()/radius() at models/pointcloud/pointnet2.py:26,28-30.
 * d_pos [n_obj,P,3].  fps: start index 0, ties -> lowest index, m samples in selection order.
 * ball query: for each centre the first `cap` (<=32) points of the same object in ascending index with
 * d2 < r2 (strict); d_nbr [n_obj,m,cap] int32 padded with -1, d_count [n_obj,m].
 * ------------------------------------------------------------------------------------------------ */
int t2p_fps(const float* d_pos, int n_obj, int P, int m, int32_t* d_idx, t2p_stream stream);
int t2p_ball_query(const float* d_pos, const int32_t* d_ctr_idx, int n_obj, int P, int m, float r2, int cap,
                   int32_t* d_nbr, int32_t* d_count, t2p_stream stream);

/* generic fused linear layer: y[M,n] (ld ldy) = act(x[M,k] (ld ldx) . wT + bias), relu in {0,1} */
int t2p_linear(const t2p_weights* w, const t2p_linear_desc* lin, const float* d_x, int M, int ldx, int relu,
               float* d_y, int ldy, t2p_stream stream);
/* rows of x[M, width] (ld) scaled to unit L2 norm, eps 1e-12 (F.normalize) */
int t2p_l2_normalize_rows(float* d_x, int M, int width, int ld, t2p_stream stream);

/* ------------------------------------------------------------------------------------------------
 * (a2,a3) PointNet2.forward(...).features2 for a batch of cells.
 * Replaces models/pointcloud/pointnet2.py:80-100 as called per cell from models/object_encoder.py:92-95.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  t2p_linear_desc sa_l1[3]; /* BN-folded first local_nn layer, k = C_in + 3 (x rows first, then pos rows) */
  t2p_linear_desc sa_l2[3]; /* BN-folded second local_nn layer */
  float sa_radius_sq[3]; /* float32(float64(r)*float64(r)), r = 0.2/0.3/0.4 (pointnet2.py:57-59) */
  t2p_linear_desc ga_l1, ga_l2; /* GlobalAbstraction mlp, k = 256 + 3 */
  t2p_linear_desc lin1, lin2;
  int32_t self_loop_quirk; /* PointConv(add_self_loops=True) flat-index self loops, see oracle/pointnet.py */
  int32_t reserved;
  int64_t ga_l2_tc_off;    /* or -1: the same images for the second global-abstraction layer (K = 512, N = 1024), one set per
                              128-wide column block: [N/128][K chunk 8][hi|lo][row n 128][64 fp16 swizzled] */
  int64_t sa_l2_tc_off[3]; /* per SA layer, or -1: fp16 hi/lo images of 2^8 * (BN-folded second layer) for the tensor-core
                              kernel (csrc/sa_tc.cu): [column block][K chunk][hi|lo][row n = output channel of the block][64
                              fp16, k = 64*chunk + e, 16-byte units XOR-swizzled by (n & 7)]; a block is always 128 output
                              channels (the UMMA M of the transposed product): sa1 (32 -> 64): one block zero-padded to 128
                              channels, K zero-padded to one chunk; sa2 (128 -> 128): one block, 2 chunks; sa3 (256 -> 256): two
                              blocks, 4 chunks each */
  int64_t dense_tc_off[6]; /* or -1: the same images (one set per 128-wide column block) for the x part of
                              the dense layers: [0] sa2 first layer (64 -> 128), [1] sa3 first layer (128 -> 256), [2] first
                              global-abstraction layer (256 -> 512), [3] lin1 (1024 -> 512), [4] lin2 (512 -> 256), [5] unused.
                              The 3-wide position part of [0]..[2] stays fp32 (added in the epilogue from the [K, N] matrix) */
} t2p_pointnet2_desc;

size_t t2p_pointnet2_workspace(const t2p_pointnet2_desc* desc, int n_obj, int P);
/* d_pos,d_rgb [n_obj,P,3]; d_obj_cell_start [n_obj]: index of the first object of the object's cell
 * (only read when self_loop_quirk); d_features2 [n_obj,256].
 * Optional debug outputs (may be NULL): d_dbg_idx[3] -> [n_obj,m_l] fps indices per layer,
 * d_dbg_nbr[3] -> [n_obj,m_l,32] int32, d_dbg_cnt[3] -> [n_obj,m_l] int32, d_dbg_x[3] -> [n_obj,m_l,C_l]. */
int t2p_pointnet2_forward(const t2p_weights* w, const t2p_pointnet2_desc* desc, const float* d_pos,
                          const float* d_rgb, const int32_t* d_obj_cell_start, int n_obj, int P,
                          float* d_features2, int32_t* const* d_dbg_idx, int32_t* const* d_dbg_nbr,
                          int32_t* const* d_dbg_cnt, float* const* d_dbg_x, void* d_ws, size_t ws_bytes,
                          t2p_stream stream);

/* ------------------------------------------------------------------------------------------------
 * (a4) ObjectEncoder.forward tail: features2 (+ colour / centre) -> object embedding.
 * Replaces models/object_encoder.py:98-138 (default flags).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  t2p_linear_desc mlp_pointnet;   /* 256 -> D */
  t2p_linear_desc color_l1, color_l2; /* 3 -> 64 -> D */
  t2p_linear_desc pos_l1, pos_l2;     /* 3 -> 64 -> D */
  t2p_linear_desc merge;              /* 3D -> D, input order (class, color, position) */
  int32_t embed_dim;
} t2p_objenc_desc;

size_t t2p_object_embed_workspace(const t2p_objenc_desc* desc, int n_obj);
int t2p_object_embed(const t2p_weights* w, const t2p_objenc_desc* desc, const float* d_features2,
                     const float* d_centers, const float* d_mean_rgb, int n_obj, float* d_emb, void* d_ws,
                     size_t ws_bytes, t2p_stream stream);

/* ------------------------------------------------------------------------------------------------
 * (a5) CellRetrievalNetwork.encode_objects tail: normalise, DynamicEdgeConv(k=8,max), global max pool,
 * lin, normalise.  Replaces models/cell_retrieval.py:94-105.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  t2p_linear_desc edge_ab; /* D -> 2D: [x_i.(W1a-W1b)+b1 | x_j.W1b] of graph1.nn layer 0 (BN folded) */
  t2p_linear_desc edge_l2; /* D -> D graph1.nn layer 1 (BN folded) */
  t2p_linear_desc lin_l1, lin_l2;
  int32_t embed_dim;
} t2p_cellagg_desc;

size_t t2p_cell_aggregate_workspace(const t2p_cellagg_desc* desc, int n_obj, int n_cells);
/* d_emb [n_obj,D] (un-normalised ObjectEncoder output), d_cell_offsets [n_cells+1] int32 (device),
 * max_cell_objects = largest object count of any cell in the call (host knows it; <= 128),
 * d_out [n_cells,D]; optional d_dbg_knn [n_obj,8] int32 (global object index or -1). */
int t2p_cell_aggregate(const t2p_weights* w, const t2p_cellagg_desc* desc, const float* d_emb,
                       const int32_t* d_cell_offsets, int n_obj, int n_cells, int max_cell_objects, float* d_out,
                       int32_t* d_dbg_knn, void* d_ws, size_t ws_bytes, t2p_stream stream);
/* the kNN of DynamicEdgeConv alone (bit-exact index parity): d_e [n_obj,D] -> d_knn [n_obj,8] */
int t2p_knn_cells(const float* d_e, const int32_t* d_cell_offsets, int n_obj, int n_cells, int max_cell_objects,
                  int D, int32_t* d_knn, int32_t* d_obj_cell, t2p_stream stream);

/* ------------------------------------------------------------------------------------------------
 * (a6) LanguageEncoder.forward (+ F.normalize of encode_text).
 * Replaces models/modules.py:74-92: Embedding + packed 1-layer biLSTM + mean of the final states.
 * The input projection is folded per vocabulary entry: xproj [2,V,4H] = emb . W_ih^T + b_ih + b_hh
 * (gate order i,f,g,o); whh [2,H,4H] = W_hh^T per direction; whh_reg = the same matrix tiled per thread.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t xproj_off;   /* [2, V, 4H] */
  int64_t whh_off;     /* [2, H, 4H] */
  int64_t whh_reg_off; /* register-resident tiling of W_hh for H in {32,64,128,256} (csrc/lstm.cu), or -1:
                          [2, H/32 (cluster rank r), H/32 (i), 4 (e), 256 (thread), 4 (gate g)] with
                          thread = w*32 + kp*4 + jj  ->  W_hh[dir][g*H + 32 r + 4 w + jj][4*(8 i + kp) + e] */
  int64_t xproj4_off;  /* H == 256 only, else -1: [2, V, H, 4] = xproj with the 4 gates of a unit contiguous */
  int64_t whh_tc_off;  /* H == 256 only, else -1: tensor-memory images of the fp16 hi/lo split of 2^8 * W_hh for the
                          tensor-core kernel (csrc/lstm_tc.cu): [2, 8 (cluster rank), 2 (hi|lo), 32 (k-unit), 128 (row
                          m = 4*unit + gate), 8 fp16 (k = 8*k-unit + e)] */
  int32_t vocab;       /* V (index 0 = <unk>/padding) */
  int32_t hidden;      /* H */
  int32_t path;        /* 0 = auto (tensor-core kernel when H == 256, else register kernel, else cluster kernel);
                          1 = shared-memory cluster kernel, 2 = register-resident kernel, 3 = tensor-core kernel */
  int32_t max_groups;  /* tensor-core kernel: clusters per direction a batch is spread over, 1..7; 0 = 7 (lowest latency).
                          A cluster takes up to 32 sequences (two ping-pong groups of <= 16); a pipelined server uses 2
                          clusters per direction for 64 queries: a little more latency per batch, about half the SM-time */
} t2p_lstm_desc;

/* Host-side tokeniser with the reference's rules (models/modules.py:60-72): '.' and ',' removed, lower-cased, split on
 * whitespace, out-of-vocabulary words -> 0; rows zero padded to max_tokens.  `texts` holds n_texts NUL-terminated UTF-8
 * strings back to back (total_bytes including terminators); h_tokens [n_texts, max_tokens] / h_lengths [n_texts] are HOST
 * buffers (typically the pinned staging buffers of the H2D copy).  ASCII rules only: callers tokenise non-ASCII strings
 * with the Unicode-aware host language instead. */
typedef struct t2p_vocab t2p_vocab;
int t2p_vocab_create(const char* const* words, const int32_t* ids, int n, t2p_vocab** out);
int t2p_vocab_destroy(t2p_vocab* v);
int t2p_tokenize(const t2p_vocab* v, const char* texts, size_t total_bytes, int n_texts, int max_tokens, int32_t* h_tokens,
                 int32_t* h_lengths, int32_t* out_max_len);

/* The same rules on the GPU (one CTA per description): the serving path ships the raw bytes of a batch and the token ids
 * never exist on the host.  t2p_vocab_to_device uploads the hash table once (current device; not during stream capture).
 * t2p_stage_texts lays the batch out in a HOST (pinned) staging buffer as
 *   [int32 byte offsets [n_texts + 1] | padding to 16 bytes | the NUL-terminated strings back to back]
 * (*used_bytes = how much of it the device needs; *all_ascii = 0 if any byte >= 0x80: such batches must take the
 * Unicode-aware host tokeniser) and, when d_stage != NULL and the batch is ASCII, enqueues that one H2D copy on `stream`.  t2p_tokenize_device reads the device copy of that buffer and writes d_tokens
 * [n_texts, max_tokens] (zero padded) and d_lengths [n_texts]; a description with more than max_tokens tokens gets length
 * max_tokens + 1, one longer than 8192 bytes gets -1 (the host checks the lengths it copies back). */
int t2p_vocab_to_device(t2p_vocab* v);
size_t t2p_stage_texts_capacity(int n_texts, size_t text_bytes);
int t2p_stage_texts(const char* texts, size_t total_bytes, int n_texts, void* h_stage, size_t stage_capacity, void* d_stage,
                    t2p_stream stream, size_t* used_bytes, int* all_ascii);
int t2p_tokenize_device(const t2p_vocab* v, const void* d_stage, int n_texts, int max_tokens, int32_t* d_tokens,
                        int32_t* d_lengths, t2p_stream stream);

/* Host fast path of the serving engine: t2p_stage_texts (+ its H2D copy), cudaGraphLaunch of the captured step
 * (`graph_exec` = cudaGraphExec_t), the D2H copy of the `out_bytes` result bytes into pinned memory and cudaEventRecord
 * (`event` = cudaEvent_t or NULL), all on `stream`, in one call.  *all_ascii = 0: the batch has non-ASCII bytes and NOTHING was
 * enqueued (the caller tokenises with the Unicode-aware host rules instead). */
int t2p_serving_submit(const char* texts, size_t total_bytes, int n_texts, void* h_stage, size_t stage_capacity, void* d_stage,
                       void* graph_exec, const void* d_out, void* h_out, size_t out_bytes, void* event, t2p_stream stream,
                       size_t* used_bytes, int* all_ascii);

/* Device-resident serving loop: launches n captured steps back to back, step i = graph_execs[i] (cudaGraphExec_t) on
 * streams[i], with no interpreter between the launches (a Python-level replay costs ~25 us per step, more than the GPU needs
 * for one).  The inputs of the captured steps (staged text, DB) are already resident; nothing is copied.  fork_join != 0: the
 * distinct streams first wait for the work enqueued on `origin` so far and `origin` waits for all of them at the end (events
 * recorded on `origin` around the call then bracket the whole region). */
int t2p_serving_replay_many(void* const* graph_execs, const t2p_stream* streams, int n, t2p_stream origin, int fork_join);

size_t t2p_lstm_encode_workspace(int B, int H);
/* d_tokens [B,T] int32 (row b valid for t < d_lengths[b]), 1 <= lengths <= T.  d_out [B,H] =
 * 0.5*(h_fwd_final + h_bwd_final), L2-normalised per row if normalize != 0. */
int t2p_lstm_encode(const t2p_weights* w, const t2p_lstm_desc* desc, const int32_t* d_tokens,
                    const int32_t* d_lengths, int B, int T, int normalize, float* d_out, void* d_ws,
                    size_t ws_bytes, t2p_stream stream);

/* ------------------------------------------------------------------------------------------------
 * (a8-a10) SuperGlue.forward: attentional GNN + final projection + log-space Sinkhorn + matching.
 * Replaces models/superglue.py:239-330 (eval-mode BatchNorm folded into mlp0).
 * Per layer (all [K,N] transposed): q,k,v,merge D->D, mlp0 2D->2D (BN folded), mlp3 2D->D.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  t2p_linear_desc q[T2P_MAX_GNN_LAYERS], k[T2P_MAX_GNN_LAYERS], v[T2P_MAX_GNN_LAYERS], merge[T2P_MAX_GNN_LAYERS];
  t2p_linear_desc mlp0[T2P_MAX_GNN_LAYERS], mlp3[T2P_MAX_GNN_LAYERS];
  int32_t is_cross[T2P_MAX_GNN_LAYERS];
  t2p_linear_desc final_proj;
  int32_t num_gnn_layers; /* len(GNN_layers) = 2 * args.num_layers */
  int32_t dim;            /* D, divisible by 4 heads */
  int32_t sinkhorn_iters;
  float bin_score;
  float match_threshold;
  int64_t tc_w_off; /* D == 128 only, else -1: fp16 hi/lo images of 2^8 * W for the tensor-core kernel (csrc/superglue_tc.cu), one
                       32 KB stage [hi|lo][128 output channels][64 fp16 of one K chunk, 16-byte units XOR-swizzled by (n & 7)]
                       after the other in consumption order: per layer q', k', v' (columns in head-major order), merge' (rows in
                       head-major order) 2 stages each; mlp0 (K = N = 256) 8 stages: column blocks 0,1 x K chunks 0,1 then column
                       blocks 0,1 x K chunks 2,3; mlp3 4 stages; after the last layer final_proj 2 stages */
  int64_t tc_b_off; /* fp32 biases for it: per layer [bq' 128 | bk' 128 | bv' 128 | bmerge 128 | b0 256 | b3 128], then bfinal 128 */
} t2p_superglue_desc;

size_t t2p_superglue_workspace(int B, int M, int N, int D);
/* With D == 128, M, N <= 32, tc_w_off >= 0 and a workspace (t2p_superglue_workspace bytes) the head runs on the tcgen05 tensor
 * cores, floor(128 / (M + N)) samples per CTA; otherwise (or if an activation leaves the fp16 range: device-side flag) on the
 * exact-fp32 CUDA-core kernel, one CTA per sample.
 * d_desc0 [B,M,D], d_desc1 [B,N,D] row layout (= reference desc.transpose(1,2)); outputs: d_P [B,M+1,N+1]
 * (exp of the log assignment), d_matches0 [B,M] / d_matches1 [B,N] int64 (-1 = unmatched),
 * d_mscores0 [B,M], d_mscores1 [B,N]; optional d_dbg_scores [B,M,N] (pre-Sinkhorn scores). */
int t2p_superglue_forward(const t2p_weights* w, const t2p_superglue_desc* desc, const float* d_desc0,
                          const float* d_desc1, int B, int M, int N, float* d_P, int64_t* d_matches0,
                          int64_t* d_matches1, float* d_mscores0, float* d_mscores1, float* d_dbg_scores,
                          void* d_ws, size_t ws_bytes, t2p_stream stream);

/* Gather variant for the cached fine stage (models/superglue_matcher.py:101-103 computes the object encodings
 * of a cell anew for every query that retrieved it -- they are query independent): sample b reads block d_idx0[b] of a
 * resident table d_desc0 [n_blocks0, M, D] and block d_idx1[b] of d_desc1 [n_blocks1, N, D] (NULL index = b). */
int t2p_superglue_forward_gather(const t2p_weights* w, const t2p_superglue_desc* desc, const float* d_desc0,
                                 const int64_t* d_idx0, const float* d_desc1, const int64_t* d_idx1, int B, int M, int N,
                                 float* d_P, int64_t* d_matches0, int64_t* d_matches1, float* d_mscores0, float* d_mscores1,
                                 float* d_dbg_scores, void* d_ws, size_t ws_bytes, t2p_stream stream);

/* ------------------------------------------------------------------------------------------------
 * (a1 / SURVEY 8f rank 2) batch_object_points on the device.  Replaces dataloading/kitti360pose/utils.py:89-110 with the
 * transforms of evaluation/pipeline.py:290-293 (FixedPoints(P) with replacement + NormalizeScale) and the per-object
 * np.mean calls of models/object_encoder.py:121-131.
 * Packed raw cell store: d_raw_xyz / d_raw_rgb [total_points, 3] float32, d_obj_offsets [n_obj + 1] int64 (points of
 * object o = [off[o], off[o+1])).  Sampling indices: d_choice [n_obj, P] int32 if given, else
 * t2p_fixed_points_index(seed, obj_id_base + o, i, P, n_o) (splitmix64 counter hash; host-callable for the oracle).
 * Outputs: d_pos / d_rgb [n_obj, P, 3] (pos centred on its float32 mean, scaled by float32((1/max|pos|) * 0.999999)),
 * d_centers / d_mean_rgb [n_obj, 3] float32 = float64 means over the RAW points; optional d_centers64 [n_obj, 3] (the pose head
 * reads float64 centres like the numpy original), optional d_choice_out [n_obj, P].
 * ------------------------------------------------------------------------------------------------ */
uint32_t t2p_fixed_points_index(uint64_t seed, uint64_t obj, uint32_t i, uint32_t P, uint32_t n);
int t2p_batch_object_points(const float* d_raw_xyz, const float* d_raw_rgb, const int64_t* d_obj_offsets, int n_obj, int P,
                            const int32_t* d_choice, uint64_t seed, int64_t obj_id_base, float* d_pos, float* d_rgb,
                            float* d_centers, float* d_mean_rgb, double* d_centers64, int32_t* d_choice_out, t2p_stream stream);

/* ------------------------------------------------------------------------------------------------
 *  pose head + accuracies on the device.
 * t2p_pose_head replaces get_pos_in_cell (models/superglue_matcher.py:138-161) for B = Q*K (query, retrieved cell) samples:
 *   d_matches0 [B, M] int64 (-1 = unmatched), d_offsets [n_off, N, 2] float32 with optional row index d_off_idx [B] (NULL: b),
 *   d_centers [n_ctr, M, 2] float64 object centres (x, y) with optional block index d_cell_idx [B] (NULL: b)
 *   -> d_pos_mean / d_pos_offsets [B, 2] float64 = mean over matched objects of centre (+ offset of its hint), (0.5, 0.5) if
 *   nothing matched; d_confidence [B] int32 = number of matched objects (evaluation/pipeline.py:196).
 * t2p_pose_accuracy replaces calc_sample_accuracies (evaluation/utils.py:31-54) and the mean-conf variant
 *   (evaluation/pipeline.py:255-263): d_cell_idx [Q, K] indexes d_cell_origin [n_cells, 2] / d_cell_size [n_cells] (float64)
 *   and d_cell_scene [n_cells]; d_pose_w [Q, 2] float64, d_pose_scene [Q]; h_top_k / h_threshs are HOST arrays (<= 8 each).
 *   d_hits [3, Q, n_k, n_t] int32: [0] in-cell mean, [1] mean with offsets, [2] mean of the most confident cell (row ik = 0).
 * ------------------------------------------------------------------------------------------------ */
int t2p_pose_head(const int64_t* d_matches0, int B, int M, int N, const float* d_offsets, const int64_t* d_off_idx,
                  const double* d_centers, const int64_t* d_cell_idx, double* d_pos_mean, double* d_pos_offsets,
                  int32_t* d_confidence, t2p_stream stream);
int t2p_pose_accuracy(const double* d_pos_mean, const double* d_pos_offsets, const int32_t* d_confidence,
                      const int64_t* d_cell_idx, int Q, int K, const double* d_cell_origin, const double* d_cell_size,
                      const int32_t* d_cell_scene, const double* d_pose_w, const int32_t* d_pose_scene, const int32_t* h_top_k,
                      int n_k, const double* h_threshs, int n_t, int32_t* d_hits, t2p_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* TEXT2POS_B200_H */
