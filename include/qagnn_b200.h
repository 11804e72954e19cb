/*
 * qagnn_b200 — C ABI of the B200-native QA-GNN message-passing path.
 *
 * The reference (michiyasunaga/qagnn) has no FFI layer: its operator API for this path is the
 * Python nn.Module surface in modeling/modeling_qagnn.py.  This header is the boundary a
 * maintainer binds (ctypes stub in INTEGRATION.md) to replace, one-for-one:
 *
 *   qagnn_graph_prep          <- the index work of GATConvE.forward     modeling_qagnn.py:419-438
 *                                + the out-degree count of message()     modeling_qagnn.py:476-479
 *   qagnn_fold_weights        <- edge_encoder on one-hots :30,:433; BatchNorm(eval) :30,:408;
 *                                linear_key/linear_msg/linear_query weight views :401-403,:464-466
 *   qagnn_gatconve_forward    <- GATConvE.forward / GATConvE.message      modeling_qagnn.py:411-484
 *                                (+ torch_geometric propagate/softmax, torch_scatter scatter)
 *   qagnn_node_feature_extra  <- QAGNN_Message_Passing.forward prologue   modeling_qagnn.py:62-73,86
 *   qagnn_mp_forward          <- QAGNN_Message_Passing.forward            modeling_qagnn.py:53-95
 *                                (mp_helper :45-50, Vh/Vx epilogue :92)
 *   qagnn_mp_core_forward /   <- the autograd graph of GATConvE.message + propagate in TRAINING mode
 *   qagnn_mp_core_backward       (modeling_qagnn.py:442,455-484 under qagnn.py:249-278 loss.backward()):
 *                                message passing on node-level projections with the attention kept for backward,
 *                                and its gradient w.r.t. the projections and the two edge tables
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns every
 *     buffer (inputs, outputs, workspaces); the library never allocates, frees or retains them;
 *   - sizes of the opaque workspaces come from the *_bytes() queries; buffers must be 256-byte
 *     aligned (torch allocations are);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); calls are re-entrant per
 *     stream and asynchronous, except qagnn_graph_prep with validate != 0, which synchronises the
 *     stream once to report out-of-range indices;
 *   - return value: 0 = ok, negative = error (qagnn_status_string).  No C++ exceptions cross
 *     the boundary;
 *   - dtypes: features fp32 row-major, indices int64 exactly as the reference's loader
 *     (utils/data_utils.py:79-197) produces them.  qagnn_gatconve_forward / qagnn_mp_forward are the eval-mode forward
 *     (dropout = identity, BatchNorm running statistics folded), the reference's evaluate_accuracy path (qagnn.py:30-38);
 *     training mode composes qagnn_mp_core_forward / _backward with the caller's dense layers (qagnn_b200/training.py).
 */
#ifndef QAGNN_B200_H_
#define QAGNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QAGNN_ABI_VERSION 3 /* 3: + training core, packed graph prep, decoder head / tail (round 2) */

enum {
  QAGNN_OK = 0,
  QAGNN_ERR_INVALID_ARGUMENT = -1, /* null pointer, non-positive size, D % H != 0, ... */
  QAGNN_ERR_CUDA = -2,             /* a launch or runtime call failed; see qagnn_last_cuda_error */
  QAGNN_ERR_INDEX_RANGE = -3,      /* edge_index / edge_type / node_type outside its range */
  QAGNN_ERR_WORKSPACE = -4,        /* caller-provided workspace is too small */
  QAGNN_ERR_UNSUPPORTED = -5       /* shape outside what the kernels are instantiated for */
};

/* Problem shape.  N nodes in the batched graph, E real directed edges (self loops excluded; the
 * library appends the N self loops itself, modeling_qagnn.py:436-438), D = emb_dim, H = head_count
 * (modeling_qagnn.py:387), T = n_ntype, R = n_etype, k = number of GATConvE layers,
 * n_per_graph = nodes per sub-graph when the batch is made of equal-sized sub-graphs laid out
 * back to back as LM_QAGNN.batch_graph does (modeling_qagnn.py:244-251), else 0 (unknown). */
typedef struct qagnn_shape {
  int64_t N;
  int64_t E;
  int32_t D;
  int32_t H;
  int32_t T;
  int32_t R;
  int32_t k;
  int32_t n_per_graph;
} qagnn_shape;

/* Raw parameters, in the reference's own layout (row-major [out, in], fp32), i.e. pointers into
 * the tensors of the module's state_dict. */
typedef struct qagnn_edge_encoder_params { /* modeling_qagnn.py:30 */
  const float *lin0_w, *lin0_b;            /* [D, R+1+2T], [D]  edge_encoder.0 */
  const float *bn_w, *bn_b, *bn_mean, *bn_var; /* [D] each     edge_encoder.1 */
  const float *lin3_w, *lin3_b;            /* [D, D], [D]      edge_encoder.3 */
} qagnn_edge_encoder_params;

typedef struct qagnn_layer_params { /* GATConvE, modeling_qagnn.py:401-408 */
  const float *key_w, *key_b;       /* [D, 3D], [D] linear_key   */
  const float *msg_w, *msg_b;       /* [D, 3D], [D] linear_msg   */
  const float *query_w, *query_b;   /* [D, 2D], [D] linear_query */
  const float *mlp0_w, *mlp0_b;     /* [D, D], [D]  mlp.0 */
  const float *bn_w, *bn_b, *bn_mean, *bn_var; /* [D] each mlp.1 */
  const float *mlp3_w, *mlp3_b;     /* [D, D], [D]  mlp.3 */
} qagnn_layer_params;

typedef struct qagnn_mp_params { /* QAGNN_Message_Passing, modeling_qagnn.py:19-38 */
  const float *emb_node_type_w, *emb_node_type_b; /* [D/2, T], [D/2]   */
  const float *emb_score_w, *emb_score_b;         /* [D/2, D/2], [D/2] */
  const float *vh_w, *vh_b, *vx_w, *vx_b;         /* [D, D], [D] each  */
  const float *score_basis;                       /* [D/2] = float32 pow(1.1, j), modeling_qagnn.py:70-71 */
} qagnn_mp_params;

int32_t qagnn_abi_version(void);
const char *qagnn_status_string(int32_t status);
/* cudaError_t of the most recent QAGNN_ERR_CUDA on this thread, as text. */
const char *qagnn_last_cuda_error(void);

/* ---- graph prep (once per forward, shared by all k layers) ------------------------------- */

/* Byte offsets of the arrays inside the opaque prep workspace (for inspection / tests).
 * E' = E + N.  All arrays are int32. */
typedef struct qagnn_prep_layout {
  size_t total_bytes;
  size_t src;          /* [E'] source of edge e (self loops at e >= E)                */
  size_t tgt;          /* [E'] target of edge e                                       */
  size_t combo;        /* [E'] (etype*T + type[src])*T + type[tgt]; self loop of v: R*T*T + type[v] */
  size_t rowptr_src;   /* [N+1] CSR by source; out-degree = rowptr_src[v+1]-rowptr_src[v] */
  size_t rowptr_tgt;   /* [N+1] CSR by target                                         */
  size_t perm_src;     /* [E'] edge ids stably sorted by source                        */
  size_t perm_tgt;     /* [E'] edge ids stably sorted by target                        */
  size_t csr_src_tgt;  /* [E'] tgt[perm_src[p]]                                       */
  size_t csr_src_combo;/* [E'] combo[perm_src[p]]                                     */
  size_t csr_tgt_src;  /* [E'] src[perm_tgt[p]]                                       */
  size_t csr_tgt_combo;/* [E'] combo[perm_tgt[p]]                                     */
  size_t csr_tgt_apos; /* [E'] position of edge perm_tgt[p] in the by-source order      */
  size_t pk_src;       /* [E'] (tgt - graph_base) << 16 | combo, by-source order (n_per_graph > 0 only) */
  size_t pk_tgt;       /* [E'] (src - graph_base) << 16 | combo, by-target order (n_per_graph > 0 only) */
  size_t csr_src_tpos; /* [E'] position of edge perm_src[p] in the by-target order       */
  size_t order_src;    /* [N]  per sub-graph: local node ids sorted by out-degree, descending (n_per_graph > 0) */
  size_t order_tgt;    /* [N]  same by in-degree: the tiled kernel gives each warp 4 nodes of similar degree   */
  size_t ninfo_src;    /* [N] x 2 words, per sub-graph in order_src order: {local id | min(out-degree, 65535) << 16, rowptr_src[v]} */
  size_t ninfo_tgt;    /* [N] x 2 words, same in order_tgt order with the in-degree and rowptr_tgt[v] (n_per_graph > 0)   */
  size_t status;       /* [4]  device-side error word + counters                        */
  size_t scratch;      /* internal                                                     */
} qagnn_prep_layout;

int32_t qagnn_graph_prep_layout(int64_t N, int64_t E, qagnn_prep_layout *out_host);
size_t qagnn_graph_prep_bytes(int64_t N, int64_t E);

/* edge_index int64 [2,E] (row 0 = source, row 1 = target), edge_type int64 [E] in [0,R),
 * node_type int64 [N] in [0,T).  validate != 0: synchronise `stream` and return
 * QAGNN_ERR_INDEX_RANGE if any index is out of range (the kernels clamp, never fault).
 * shape->n_per_graph > 0 additionally asserts that no edge crosses a sub-graph boundary
 * (src / n_per_graph == tgt / n_per_graph), which LM_QAGNN.batch_graph guarantees; a violation is
 * reported as QAGNN_ERR_INDEX_RANGE as well.  It enables the shared-memory-tiled kernels. */
int32_t qagnn_graph_prep(const int64_t *edge_index, const int64_t *edge_type, const int64_t *node_type,
                         const qagnn_shape *shape, void *prep, size_t prep_bytes, int32_t validate,
                         void *stream);

/* The same prep for a PACKED batch: the edges of sub-graph g occupy [graph_ptr[g], graph_ptr[g+1]) of edge_index / edge_type (what
 * LM_QAGNN.batch_graph, modeling_qagnn.py:244-251, and qagnn_b200.data.pack_adj produce; graph_ptr int64 [N/n_per_graph + 1] on the
 * DEVICE) and max_edges_per_graph bounds their length (host value).  One CTA builds everything for one sub-graph in shared memory:
 * one launch instead of twelve, bit-identical arrays.  QAGNN_ERR_UNSUPPORTED when a sub-graph does not fit one CTA's shared
 * memory (use qagnn_graph_prep); an edge outside its sub-graph's node range sets the same status bit as qagnn_graph_prep. */
int32_t qagnn_graph_prep_packed(const int64_t *edge_index, const int64_t *edge_type, const int64_t *node_type,
                                const int64_t *graph_ptr, int32_t max_edges_per_graph, const qagnn_shape *shape, void *prep,
                                size_t prep_bytes, int32_t validate, void *stream);

/* ---- weight folding (once per set of weights) --------------------------------------------- */

size_t qagnn_fold_bytes(const qagnn_shape *shape);
/* `layers_host` is a HOST array of shape->k structs holding DEVICE pointers; `mp_host` may be NULL
 * when only qagnn_gatconve_forward will be used. */
int32_t qagnn_fold_weights(const qagnn_shape *shape, const qagnn_edge_encoder_params *edge_encoder_host,
                           const qagnn_layer_params *layers_host, const qagnn_mp_params *mp_host,
                           void *folded, size_t folded_bytes, void *stream);

/* ---- forward ----------------------------------------------------------------------------------- */

size_t qagnn_forward_workspace_bytes(const qagnn_shape *shape);

/* One GATConvE layer `layer` (0-based index into the folded blob):
 *   out[N,D] = mlp(propagate(...))   (NO GELU: that is applied by mp_helper, modeling_qagnn.py:48)
 * alpha_out (optional, may be NULL): [E+N, H] softmax weights BEFORE the out-degree rescale, in
 * edge_index' order (real edges, then self loops) — what return_attention_weights=True returns.
 * aggr_out (optional): [N,D] propagate() output before the node MLP. */
int32_t qagnn_gatconve_forward(const qagnn_shape *shape, int32_t layer, const float *x, const float *extra,
                               const void *prep, const void *folded, float *out, float *alpha_out,
                               float *aggr_out, void *workspace, size_t workspace_bytes, void *stream);

/* node_feature_extra[N,D] = [GELU(emb_node_type(onehot(type))) ‖ GELU(emb_score(sin(basis*score)))] */
int32_t qagnn_node_feature_extra(const qagnn_shape *shape, const int64_t *node_type, const float *node_score,
                                 const void *folded, float *extra_out, void *workspace, size_t workspace_bytes,
                                 void *stream);

/* Whole QAGNN_Message_Passing.forward (eval): H_in [N,D], node_type int64 [N], node_score [N],
 * out [N,D] = GELU(Vh(H_in) + Vx(X_k)).  x_layers_out (optional): [k,N,D] activations after each
 * layer's GELU. */
int32_t qagnn_mp_forward(const qagnn_shape *shape, const float *H_in, const int64_t *node_type,
                         const float *node_score, const void *prep, const void *folded, float *out,
                         float *x_layers_out, void *workspace, size_t workspace_bytes, void *stream);

/* ---- training --------------------------------------------------------------- */

/* Message passing of one layer on caller-computed node-level projections, keeping what backward needs.
 *   qkm [N, 3D] row-major = [Q/sqrt(d) | Kx | Mx]  (Q = linear_query([x|extra])/sqrt(d), Kx / Mx = the [x|extra] part of
 *   linear_key / linear_msg), ke / me [C, D] = edge-table part of linear_key / linear_msg (+ bias) per combo,
 *   C = R*T*T + T (see qagnn_prep_layout.combo).
 * Outputs: aggr [N, D] (propagate() output), alpha_scaled [E+N, H] = softmax * out-degree in BY-SOURCE order (saved for
 * backward), alpha_out (optional) [E+N, H] un-scaled softmax in edge_index' order.  `scratch` = [E+N, H] floats. */
int32_t qagnn_mp_core_forward(const qagnn_shape *shape, const void *prep, const float *qkm, const float *ke, const float *me,
                              float *aggr, float *alpha_scaled, float *alpha_out, float *scratch, void *stream);

/* Gradient of qagnn_mp_core_forward: given d_aggr [N, D] writes d_qkm [N, 3D], d_ke / d_me [C, D] (zeroed here first).
 * combo_order int32 [E+N]: the by-source edge positions stably sorted by qagnn_prep_layout.csr_src_combo (the caller
 * builds it once per batch).  `scratch` = [E+N, H] floats.  Only dKe / dMe use atomics (a few thousand vector adds). */
int32_t qagnn_mp_core_backward(const qagnn_shape *shape, const void *prep, const int32_t *combo_order, const float *qkm,
                               const float *ke, const float *me, const float *alpha_scaled, const float *d_aggr,
                               float *d_qkm, float *d_ke, float *d_me, float *scratch, void *stream);

/* A stand-alone dense layer on the same tensor-core path the forward uses (diagnostics / tests):
 *   C[M,N] (ldc) = act( [A1 | A2] @ W^T + bias ),  fp32 row-major operands, W [N, K1+K2] (ldw), act 0/1/2 = none/ReLU/GELU(tanh).
 * Splits the operands into bf16 hi/lo planes in `workspace` and runs the tcgen05 3-pass GEMM; returns
 * QAGNN_ERR_UNSUPPORTED when the tensor-core path cannot take the shape (K or ld not multiples of 8). */
size_t qagnn_linear_workspace_bytes(int64_t M, int32_t N, int32_t K1, int32_t K2);
int32_t qagnn_linear_bf16x3(const float *A1, int32_t lda1, int32_t K1, const float *A2, int32_t lda2, int32_t K2,
                            const float *W, int32_t ldw, const float *bias, float *C, int32_t ldc, int64_t M, int32_t N,
                            int32_t act, void *workspace, size_t workspace_bytes, void *stream);

/* The step after the path: QAGNN's masked multi-head attention pooling of the node representations,
 * MultiheadAttPoolLayer.forward (utils/layers.py:344-371, called at modeling_qagnn.py:180), eval mode, fused into one
 * kernel that reads X [B, n, D] once.  qs [B, D] = w_qs(sent_vecs) (computed by the caller), mask uint8 [B, n]
 * (1 = masked out), wk/wv [D, D] and bk/bv [D] = w_ks / w_vs weights; outputs pooled [B, D], attn [n_head*B, n]. */
int32_t qagnn_attention_pool(int32_t B, int32_t n, int32_t D, int32_t n_head, const float *X, const float *qs,
                             const uint8_t *mask, const float *wk, const float *bk, const float *wv, const float *bv,
                             float *pooled, float *attn, void *stream);

/* The same pooling with the rest of QAGNN.forward's tail (modeling_qagnn.py:172-187) in the kernel: the pool mask is derived
 * from adj_lengths [B] and node_type [B,n] ((i >= len) | (type == 3), node 0 kept if everything is masked) and the row of
 * cat(graph_vecs, sent_vecs, Z) is written in place into concat [B, 2D+S] (Z = X[b,0,:], sent_vecs [B,S]).  The caller
 * all-gathers `concat` (multi-GPU) and applies the answer MLP. */
int32_t qagnn_decoder_tail(int32_t B, int32_t n, int32_t D, int32_t n_head, int32_t S, const float *X, const float *qs,
                           const int64_t *node_type, const int64_t *adj_lengths, const float *sent_vecs, const float *wk,
                           const float *bk, const float *wv, const float *bv, float *concat, float *attn, void *stream);

/* The step before the path: QAGNN.forward's input assembly (modeling_qagnn.py:153-167), eval mode, one kernel:
 *   H_out[b,0,:] = ctx[b,:] (= GELU(svec2nvec(sent_vecs)), caller-computed [B,D]);  H_out[b,i,:] = table[concept_ids[b,i]-1], i >= 1,
 *   with `table` [n_concept, D] the concept embedding after cpt_transform + GELU (folded once by the caller; the table is
 *   frozen in eval);  scores_out [B,n] = the reference's relevance-score normalisation of node_scores [B,n] by adj_lengths. */
int32_t qagnn_decoder_head(int32_t B, int32_t n, int32_t D, const int64_t *concept_ids, int64_t n_concept, const float *table,
                           const float *ctx, const float *node_scores, const int64_t *adj_lengths, float *H_out,
                           float *scores_out, void *stream);

/* Launch counters since load (kernels this library enqueued); for bench.py's gpu_launches. */
int64_t qagnn_launch_count(void);

/* Optional device-side stage timing (diagnostics for bench.py's roofline; single-threaded use).
 * While enabled, every forward brackets its stages with cudaEvents on the launching stream.
 * qagnn_profile_read synchronises those events and returns, per stage, the summed milliseconds and
 * the number of timed intervals since the last enable.  Stages: */
enum {
  QAGNN_PROF_GRAPH_PREP = 0,   /* qagnn_graph_prep                                   */
  QAGNN_PROF_PROJECTION = 1,   /* [x|extra] @ Wp^T  (Q|Kx|Mx)                         */
  QAGNN_PROF_MESSAGE_PASSING = 2, /* logits + per-source softmax + rescale + per-target sum */
  QAGNN_PROF_NODE_MLP = 3,     /* mlp.0 + BN + ReLU + mlp.3 (+GELU)                    */
  QAGNN_PROF_PRO_EPILOGUE = 4, /* node_feature_extra and Vh/Vx                         */
  QAGNN_PROF_STAGES = 5
};
int32_t qagnn_profile_enable(int32_t on);
int32_t qagnn_profile_read(double *ms_out /* [QAGNN_PROF_STAGES] */, int64_t *count_out /* [QAGNN_PROF_STAGES] */);

#ifdef __cplusplus
}
#endif
#endif /* QAGNN_B200_H_ */
