/* gtb200 -- C ABI of the B200-native Interaction-Network hot path.
 *
 * Plain C, no torch types: raw DEVICE pointers, sizes, a cudaStream_t passed as
 * void*.  The caller owns every buffer (inputs, outputs, workspace); no entry
 * point allocates device memory or synchronises the stream.  Every function
 * returns 0 on success or a negative GTB_ERR_* code; gtb_last_error() gives the
 * thread-local message.
 *
 * Each entry point cites the reference code it replaces.  Citations are
 * relative to /root/reference/src/gnn_tracking (gnn-tracking/gnn_tracking @ 23.12.1).
 *
 * The reference has no FFI of its own (pure Python over torch / PyG); the
 * binding a maintainer would add is the ctypes stub shown in INTEGRATION.md and
 * implemented in gnn_tracking_b200/_lib.py.
 */
#ifndef GTB200_H
#define GTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTB_OK                    0
#define GTB_ERR_BAD_ARG          -1
#define GTB_ERR_UNSUPPORTED_DIM  -2
#define GTB_ERR_WORKSPACE        -3
#define GTB_ERR_CUDA             -4
#define GTB_ERR_ARCH             -5

#define GTB_MAX_SRCS   16   /* column blocks of a concatenated MLP input          */
#define GTB_MAX_LAYERS 3    /* Linear layers of an MLP on the path (mlp.py:44-51) */
#define GTB_MAX_WIDTH  128  /* max Linear output width of the FFMA path           */

/* final activation of a fused MLP */
#define GTB_ACT_NONE           0
#define GTB_ACT_RELU           1
#define GTB_ACT_SIGMOID_AFFINE 2 /* eps + (1-2 eps) * sigmoid(v): edge_classifier.py:115-117,
                                    track_condensation_networks.py:284-288 */

/* which implementation a fused-MLP call must use */
#define GTB_IMPL_AUTO   0
#define GTB_IMPL_FFMA   1 /* fp32 CUDA-core tiles (any width <= GTB_MAX_WIDTH)             */
#define GTB_IMPL_TCGEN05 2 /* tcgen05 3xTF32 tiles (Linear widths <= 64)     */

int         gtb_version(void);
const char* gtb_last_error(void);
/* 0 when `device` is a compute-capability 10.x part; GTB_ERR_ARCH otherwise. */
int gtb_arch_ok(int device);

/* ------------------------------------------------------------------ graph plan
 * Destination-sorted view of edge_index, built once per graph and shared by all
 * layers.  Replaces what PyG's MessagePassing.propagate re-derives on every call
 * (models/interaction_network.py:67 -> index_select on edge_index[0|1] and
 * scatter_add_ over edge_index[1]).
 *
 *  edge_index : int64 [2, E] row-major as delivered by the reference API
 *               (row 0 = source j, row 1 = target i).
 *  perm       : int32 [E]   stable argsort of edge_index[1]  (bit-exact contract)
 *  rowptr     : int32 [N+1] CSR offsets of the sorted list
 *  src_sorted : int32 [E]   edge_index[0][perm]
 *  dst_sorted : int32 [E]   edge_index[1][perm]
 *  status     : int32 [1]   set to nonzero if an index is outside [0, N)
 */
size_t gtb_plan_workspace_bytes(int64_t n_nodes, int64_t n_edges);
int gtb_plan_build(const int64_t* edge_index, int64_t n_nodes, int64_t n_edges,
                   int32_t* perm, int32_t* rowptr, int32_t* src_sorted, int32_t* dst_sorted,
                   int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* Plan of an edge sub-graph (track_condensation_networks.py:251-252,
 * Data.edge_subgraph(W > thr)): a dst-sorted list stays sorted under filtering, so
 * this is a stream compaction of the parent plan.  keep: uint8 [E] in ORIGINAL
 * edge order.  new_id: int32 [E] scratch, receives the position of every kept edge
 * in the compacted original order (-1 for dropped); kept_ids: int32 [E], the
 * inverse map (original id of compacted edge j) in its first n_kept entries.
 * n_kept_out: int32 [1]. */
size_t gtb_plan_filter_workspace_bytes(int64_t n_nodes, int64_t n_edges);
int gtb_plan_filter(const uint8_t* keep, int64_t n_nodes, int64_t n_edges,
                    const int32_t* perm, const int32_t* src_sorted, const int32_t* dst_sorted,
                    int32_t* new_id, int32_t* kept_ids, int32_t* perm_out, int32_t* rowptr_out,
                    int32_t* src_sorted_out, int32_t* dst_sorted_out, int32_t* n_kept_out,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Plan of the graph with its orphan nodes removed (track_condensation_networks.py:254-259: the unique
 * endpoints of the surviving edges, relabelled in increasing order).  The relabelling is monotone, so the
 * plan keeps its order: `perm` stays valid as it is, the endpoints are relabelled and rowptr compacted.
 * new_id: int32 [N] new id of every node (-1 for an orphan); node_ids: int32 [N], original id of new node
 * j in its first n_kept entries; rowptr_out: int32 [N + 1] (n_kept + 1 used); n_kept_out: int32 [1]. */
size_t gtb_plan_prune_workspace_bytes(int64_t n_nodes);
int gtb_plan_prune_orphans(int64_t n_nodes, int64_t n_edges, const int32_t* rowptr,
                           const int32_t* src_sorted, const int32_t* dst_sorted,
                           int32_t* new_id, int32_t* node_ids, int32_t* rowptr_out,
                           int32_t* src_sorted_out, int32_t* dst_sorted_out, int32_t* n_kept_out,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------ packed MLP weights
 * An MLP of the path (models/mlp.py:18-62: Linear/ReLU chain, nn.Linear weights
 * [out, in] row-major, optional bias) is repacked once per weight version into the
 * K-major, zero-padded layout the tiles consume.  `impl` selects the layout
 * (GTB_IMPL_FFMA or GTB_IMPL_TCGEN05).  dims = {K0, N0, N1, N2} (true widths).
 * block_widths[n_blocks] are the widths of the concatenated (non-projected) source blocks the
 * MLP will be called with (they sum to K0; NULL = one block): the tcgen05 layout pads every
 * block to a multiple of 8 columns.  gtb_mlp_packed_bytes returns 0 when `impl` does not
 * support the widths. */
size_t gtb_mlp_packed_bytes(int n_layers, const int32_t* dims, int n_blocks, const int32_t* block_widths,
                            int impl);
int gtb_mlp_pack(int n_layers, const int32_t* dims, int n_blocks, const int32_t* block_widths,
                 const float* const* weights, const float* const* biases /* entries may be NULL */,
                 int impl, void* packed, void* stream);

/* Number of 32 KB shared-memory staging slots the tcgen05 tiles keep beside the packed weights of
 * this MLP (0: widths unsupported).  The host side uses it to split a long Linear chain where the
 * weights would leave fewer than two slots (no gather in flight while a tile is computed): the W
 * head of the edge classifier, edge_classifier.py:79-84 with 2*Dn + De*(L+1) input columns. */
int gtb_mlp_tc_slots(int n_layers, const int32_t* dims, int n_blocks, const int32_t* block_widths);

/* flags of a source block.
 * GTB_SRC_PROJECTED: the block was already multiplied by its column block of the first Linear
 *   (a per-node table [*, N0], N0 = dims[1]); the gathered rows are ADDED to the output of
 *   the first Linear instead of being concatenated in front of it.  With E >> N this moves
 *   2*Dn*H of the (2*Dn + De)*H multiply-adds per edge of the relational model
 *   (interaction_network.py:86-87) to a per-node launch.  Such blocks do not count in dims[0].
 * GTB_SRC_SORTED: hint, `index` is non-decreasing (neighbouring rows repeat). */
#define GTB_SRC_PROJECTED 1
#define GTB_SRC_SORTED    2

/* --------------------------------------------------------------- fused row MLP
 * out[orow(r), :] = epilogue( MLP( cat_s( act_s( src_s[irow_s(r), 0:width_s] ) ) ) )
 * for r in [0, n_rows): the one building block behind every dense op of the path:
 *   - relational model  interaction_network.py:75-89  (3 gathered column blocks)
 *   - object model      interaction_network.py:92-103 (+ residual, resin.py:17-42)
 *   - encoders          edge_classifier.py:102-103, track_condensation_networks.py:279-280
 *   - W head            edge_classifier.py:108-117
 *   - beta / H heads    track_condensation_networks.py:284-298
 * The concatenation is never materialised. */
typedef struct {
  const float*   ptr;    /* [*, ld] fp32 rows                                           */
  const int32_t* index;  /* row gather index per output row, or NULL for identity        */
  int32_t        width;  /* columns taken (from column 0)                                */
  int32_t        ld;     /* row stride in elements                                       */
  int32_t        relu;   /* 1: relu on load (resin.py:104-105: layers > 0 see relu(x))   */
  int32_t        flags;  /* GTB_SRC_*                                                    */
} gtb_src_t;

typedef struct {
  int64_t   n_rows;
  int32_t   n_srcs;
  int32_t   n_layers;               /* 1..GTB_MAX_LAYERS                                  */
  gtb_src_t srcs[GTB_MAX_SRCS];
  int32_t   dims[GTB_MAX_LAYERS + 1]; /* K0 (= sum of src widths), N0, N1, N2              */
  const void* packed;               /* from gtb_mlp_pack (same impl)                      */
  int32_t   impl;                   /* GTB_IMPL_*                                         */
  int32_t   final_act;              /* GTB_ACT_*                                          */
  float     act_eps;
  /* out = res_b * value (+ res_a * res[r, :] when res != NULL), then * (*out_scale) when
   * out_scale != NULL.  sqconvex_combination resin.py:17-42: res_a = sqrt(alpha),
   * res_b = sqrt(1-alpha); out_scale: the learnable _latent_normalization scalar kept on
   * the device (track_condensation_networks.py:298).  res_b must be 1 for "no scaling". */
  float     res_a, res_b;
  int32_t   res_ld;
  const float* res;
  /* optional per-row scale applied to every source block on load, used for the L2 row
   * normalisation in front of ResFCNN (mlp.py:115-116); NULL = none */
  const float* row_scale;
  const float* out_scale;
  /* output rows (out may be NULL when only the aggregate is wanted) */
  float*    out;
  const int32_t* out_index;         /* row scatter index or NULL                          */
  int32_t   out_ld;
  /* optional per-destination sum of the output rows (PyG SumAggregation,
   * interaction_network.py:22,36): rows must be ordered so that seg_id is
   * non-decreasing (the plan's dst_sorted); aggr [n_segments, aggr_ld] must be
   * zero-filled by the caller. */
  int32_t   aggr_ld;
  float*    aggr;
  const int32_t* seg_id;
  const int32_t* rowptr;
  /* optional gate [n_rows, gate_ld] in launch-row order: out = gate > 0 ? value : 0 after the
   * epilogue above.  The backward pass multiplies dY W by the ReLU mask of the recomputed
   * activation with it (d relu, models/mlp.py:44-51). */
  const float* gate;
  int32_t   gate_ld;
  /* optional: the two hidden activations (post-ReLU outputs of the first and the second Linear) of a
   * 3-Linear launch, [n_rows, hidden_ld] in launch-row order -- what the backward pass otherwise recomputes
   * (models/mlp.py:59-62 under autograd keeps them as saved tensors).  Only the dedicated 64-wide kernels
   * write them: ask gtb_fused_mlp_saves_hidden first; a launch that cannot honour the request is refused. */
  int32_t   hidden_ld;
  float*    hidden0;
  float*    hidden1;
} gtb_mlp_desc_t;

int gtb_fused_mlp_f32(const gtb_mlp_desc_t* desc, void* stream);
/* 1 if gtb_fused_mlp_f32 would run this descriptor on a kernel that can write hidden0 / hidden1, else 0. */
int gtb_fused_mlp_saves_hidden(const gtb_mlp_desc_t* desc);

/* Test hook (synchronises): *flag != 0 if a tcgen05 kernel ever gave up waiting for its MMA
 * barrier -- such a kernel traps, so the CUDA error is sticky as well. */
int gtb_debug_tc_timeout(int* flag);
/* Test hook (synchronises): per-stage clock accumulation of the tcgen05 kernel by thread 0 of CTA 0.
 * out32 == NULL: set the enable flag for later launches and zero the counters; else copy the 32
 * counters out (see tests/cuda/tc_diag.py for the stage names). */
int gtb_debug_tc_profile(int enable, long long* out32);

/* ------------------------------------------------------------- IN layer wrappers
 * One Interaction-Network layer (interaction_network.py:54-103) on a planned graph.
 * edge_attr / e_tilde stay in the caller's edge order; kernels walk them in
 * dst-sorted order through `perm`.
 *   gtb_in_edge_forward_f32 : e_tilde = MLP_rel(cat[x[dst], x[src], edge_attr]) and
 *                             aggr[i] = sum_{dst(e)=i} e_tilde[e]      (aggr zeroed inside)
 *   gtb_in_node_forward_f32 : x_out = res_a * x + res_b * MLP_obj(cat[act(x), aggr])
 *                             (res_a = 0, res_b = 1 for a bare IN layer)
 */
int gtb_in_edge_forward_f32(const float* x, int32_t x_ld, int32_t relu_x,
                            const float* edge_attr, int32_t e_ld, int32_t relu_e,
                            int64_t n_nodes, int64_t n_edges,
                            const int32_t* perm, const int32_t* rowptr,
                            const int32_t* src_sorted, const int32_t* dst_sorted,
                            int32_t node_dim, int32_t edge_dim, int32_t hidden, int32_t edge_outdim,
                            const void* packed_rel, int impl,
                            float* e_tilde, int32_t eo_ld, float* aggr, void* stream);
int gtb_in_node_forward_f32(const float* x, int32_t x_ld, int32_t relu_x,
                            const float* aggr, int64_t n_nodes,
                            int32_t node_dim, int32_t aggr_dim, int32_t hidden, int32_t node_outdim,
                            const void* packed_obj, int impl,
                            float res_a, float res_b, const float* res, int32_t res_ld,
                            float* x_out, int32_t xo_ld, void* stream);

/* Edge encoder of the edge classifier in one launch (csrc/enc_ws.cu): a two-Linear MLP 4 -> 64 -> 64 over gathered
 * rows (edge_classifier.py:103 `relu(ec_edge_encoder(edge_attr))`, models/mlp.py:18-62 with L = 2):
 *   out[r] = act(W1 relu(W0 x[index ? index[r] : r] + b0) + b1),   act = ReLU iff final_relu
 * x fp32 [x_rows, x_ld] with 4 feature columns (x_rows: rows of the gathered table, 0 = unknown), index int32 [n_rows]
 * or NULL, w0 fp32 [64, 4] and b0 fp32 [64] / NULL in nn.Linear layout (the first Linear runs on the CUDA cores),
 * packed_w1: gtb_mlp_pack image (GTB_IMPL_TCGEN05) of the one Linear {64, 64}, out fp32 [n_rows, out_ld].
 * Row strides in elements, multiples of 4; pointers 16-byte aligned. */
int gtb_edge_encoder_f32(const float* x, int32_t x_ld, const int32_t* index, int64_t n_rows, int64_t x_rows, const float* w0,
                         const float* b0, const void* packed_w1, int32_t final_relu, float* out, int32_t out_ld, void* stream);

/* Node side of one 64-wide Interaction-Network layer in ONE launch (csrc/node_ws.cu): the object model with
 * the residual of the stack, the two per-node products the NEXT consumer gathers (GTB_SRC_PROJECTED blocks of
 * the next layer's relational model, interaction_network.py:75-89, or of the W head, edge_classifier.py:108-117),
 * and the aggregate handed back zeroed for the next layer's edge kernel:
 *   x_out = res_a * res + res_b * MLP_obj(cat[act(x), aggr])     (interaction_network.py:92-103, resin.py:17-42)
 *   p_a   = act'(x_out) Wa^T,  p_b = act'(x_out) Wb^T            (act' = ReLU iff proj_relu)
 *   aggr  = 0                                                     (iff zero_aggr)
 * packed_obj: gtb_mlp_pack image (GTB_IMPL_TCGEN05) of dims {128, 64, 64, 64} with blocks {64, 64}; NULL:
 * projection only, p_a / p_b of act'(x).  packed_pa / packed_pb: gtb_mlp_pack images of one bias-free Linear
 * {64, 64} each; NULL: no projections.  All tables fp32, 64 columns, row strides in elements (multiples of 4),
 * pointers 16-byte aligned; res may be NULL (no residual term) or alias x. */
int gtb_in_node_fused_f32(const float* x, int32_t x_ld, int32_t relu_x, float* aggr, int32_t aggr_ld, int32_t zero_aggr,
                          int64_t n_nodes, const void* packed_obj, float res_a, float res_b, const float* res, int32_t res_ld,
                          float* x_out, int32_t xo_ld, const void* packed_pa, const void* packed_pb, int32_t proj_relu,
                          float* p_a, int32_t pa_ld, float* p_b, int32_t pb_ld, void* stream);

/* bf16 variant of the edge kernel for the reference's mixed-precision runs (torch.autocast(bfloat16) around
 * models/interaction_network.py:75-89; BASELINE config 3: GraphTCN with node = edge = hidden width 128).
 * All feature tables are bf16 [*, ld] (ld in elements, 16-byte multiples), fp32 accumulation on the tensor
 * cores (tcgen05 kind::f16), every Linear output rounded to bf16 as autocast's Linear does, the
 * per-destination sum accumulated in fp32:
 *   e_out[o(r)] = bf16(W2 relu(bf16(W1 relu(bf16(W0e act(e_in[i(r)]) + P_i[dst(r)] + P_j[src(r)] + b0)) + b1)) + b2)
 *   aggr[dst(r), 0:128] += e_out row      (aggr fp32 [N, aggr_ld], zero-filled by the caller)
 * rows r walk the plan's destination-sorted edge list; e_index / out_index: `perm`, or NULL when the edge
 * features are already / stay in that order.  P_i / P_j: act(x) times the two node column blocks of the first
 * Linear, per node (the caller's GEMM).  gtb_in_edge_bf16_pack: weights fp32 [128, 128] x 3 (nn.Linear
 * layout; W0 = the EDGE columns of the first Linear), biases fp32 [128] x 3 (entries may be NULL). */
size_t gtb_in_edge_bf16_packed_bytes(void);
int gtb_in_edge_bf16_pack(const float* const* weights, const float* const* biases, void* packed, void* stream);
int gtb_in_edge_forward_bf16(const void* e_in, int32_t e_ld, const int32_t* e_index, int32_t relu_e,
                             const void* p_i, int32_t pi_ld, const void* p_j, int32_t pj_ld, int64_t n_edges,
                             const int32_t* src_sorted, const int32_t* dst_sorted, const void* packed,
                             void* e_out, int32_t eo_ld, const int32_t* out_index,
                             float* aggr, int32_t aggr_ld, void* stream);

/* ------------------------------------------------------------------- EC losses
 * metrics/losses/ec.py.  y_true may be NULL-free uint8 or float labels:
 *   label_kind 0: float [E], 1: uint8/bool [E].
 * If pt != NULL, labels are falsified: y &= pt[src[e]] > pt_thld (ec.py:71-92;
 * only edge_index[0] is looked at).  src: int64 [E] (edge_index row 0).
 * out: float [2] = {sum, n}; the mean is sum / n (finished on the host side of the
 * boundary so that the kernel stays allocation- and sync-free).  out must be zeroed.
 *   mode 0: BCE (ec.py:116-121, log clamped at -100 as torch does)
 *   mode 1: focal (ec.py:12-29) with alpha, gamma, scalar pos_weight
 *   mode 2: haughty focal (ec.py:153-178): falsified labels are pos_weight,
 *           the raw labels the target. */
int gtb_ec_loss_f32(const float* w, const void* y, int label_kind, int64_t n_edges,
                    const int64_t* src, const float* pt, float pt_thld,
                    int mode, float alpha, float gamma, float pos_weight,
                    double* out /* [2] */, void* stream);

/* dw[e] = (*scale) * d term_e / d w_e for the same modes (scale = upstream gradient / n_edges, a
 * device scalar): backward of the losses above; BCE as torch does, (w - y) / max(w (1 - w), 1e-12). */
int gtb_ec_loss_grad_f32(const float* w, const void* y, int label_kind, int64_t n_edges,
                         const int64_t* src, const float* pt, float pt_thld,
                         int mode, float alpha, float gamma, float pos_weight,
                         const float* scale, float* dw, void* stream);

/* ------------------------------------------------------- condensation loss (tiger)
 * metrics/losses/oc.py:251-347 without the N x K planes.
 *   beta [N], x [N, d] (ld = d), object_id int64 [N], object_mask uint8 [N]
 *   (object_mask from get_good_node_mask_tensors, utils/graph_masks.py:19-28).
 * Step 1 (host side picks K = #unique masked ids; ids are compacted on the device):
 *   gtb_oc_prepare   : sorts / uniques the masked object ids -> uniq [K<=N] int64,
 *                      obj_slot int32 [N] (slot of the hit's object or -1), n_uniq [1]
 *   gtb_oc_alphas    : alpha_k = argmax_j q_j [id_j == uniq_k]  (first index on ties)
 *   gtb_oc_potentials: the four sums; out double[8] =
 *                      {V_att_sum, V_rep_sum, coward_sum, noise_sum, n_noise, n_hits_oi, n_rep, K}
 */
size_t gtb_oc_workspace_bytes(int64_t n_nodes);
int gtb_oc_prepare(const int64_t* object_id, const uint8_t* object_mask, int64_t n_nodes,
                   int64_t* uniq, int32_t* obj_slot, int32_t* n_uniq,
                   void* workspace, size_t workspace_bytes, void* stream);
int gtb_oc_alphas(const float* beta, const int32_t* obj_slot, int64_t n_nodes, float q_min,
                  int32_t k, unsigned long long* packed_scratch /* [k] */, int32_t* alphas /* [k] */,
                  void* stream);
int gtb_oc_potentials(const float* beta, const float* x, int32_t d, const int64_t* object_id,
                      const uint8_t* object_mask, const int32_t* obj_slot, int64_t n_nodes,
                      const int32_t* alphas, int32_t k, float q_min, int64_t noise_threshold,
                      double* out /* [8], zeroed by caller */, void* stream);

/* Gradients of the four sums above w.r.t. beta [n] and x [n, d] (what torch autograd derives for
 * oc.py:282-336: cdist with a zero sub-gradient at distance 0, indexing by alphas, masked sums).
 *   coef  float[4] on the device: upstream gradient of each loss term divided by its normaliser,
 *         {attractive, repulsive, coward, noise}
 *   gq    float[n] scratch (d loss / d q_j);  gbeta float[n], gx float[n, d]: outputs (overwritten) */
int gtb_oc_potentials_grad(const float* beta, const float* x, int32_t d, const int64_t* object_id,
                           const int32_t* obj_slot, int64_t n_nodes, const int32_t* alphas, int32_t k,
                           float q_min, int64_t noise_threshold, const float* coef, float* gq,
                           float* gbeta, float* gx, void* stream);

/* ------------------------------------------------ radius-graph potentials (hinge, RG)
 * Replaces torch_cluster.radius_graph(x, r, batch, loop=False, max_num_neighbors) + the sums over
 * its edges (metrics/losses/metric_learning.py:93-112,47-52; metrics/losses/oc.py:46-69,115-117)
 * without materialising the edge list.  An edge (neighbour j -> centre i) exists when
 * ||x_i - x_j||^2 < r^2, i != j, batch[i] == batch[j] (batch may be NULL); a centre keeps its
 * max_num_neighbors lowest-index neighbours.  Kept edges additionally need src_flag[j] != 0 and
 * pid[j] != pid[i].
 *   mode 0: term = relu(r - dist^p)                         (hinge repulsion; src_flag = hits of interest)
 *   mode 1: term = (r - sqrt(eps + dist^2)) * q_j * q_i     (condensation repulsion; src_flag = is
 *           condensation point; q = atanh(beta)^2 + q_min; beta required)
 * out double[4] += {sum of terms, kept edges, sum of beta over pid == 0, hits with pid == 0}
 * (the last two only when beta != NULL).  x: [n, d] fp32, d <= 16. */
int gtb_radius_pair_sum_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, const int64_t* pid,
                            const uint8_t* src_flag, const float* beta, float q_min, float r, float p,
                            float eps, int32_t max_num_neighbors, int32_t mode, double* out, void* stream);
/* out double[2] += {sum over edges e with src_flag[edges[0][e]] of ||x_a - x_b||^p, their number}:
 * attractive hinge term (metric_learning.py:28-30,111).  edges: int64 [2, n_edges]. */
int gtb_edge_dist_pow_sum_f32(const float* x, int32_t d, const int64_t* edges, int64_t n_edges,
                              const uint8_t* src_flag, float p, double* out, void* stream);

/* torch_cluster.radius_graph(x, r, batch, loop, max_num_neighbors) as an edge list (the un-vendored
 * native op behind metrics/losses/oc.py:115-117, metric_learning.py:97-103, 232): strict dist < r, same
 * batch entry, at most max_num_neighbors neighbours per centre (the lowest indices).  Two passes:
 *   count: counts int32 [n] = kept neighbours per centre;
 *   fill : offsets int64 [n] = exclusive prefix sum of counts, edge_index int64 [2, n_edges] with
 *          row 0 = neighbour, row 1 = centre, grouped by centre, neighbours ascending. */
int gtb_radius_graph_count_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, float r,
                               int32_t max_num_neighbors, int32_t loop, int32_t* counts, void* stream);
int gtb_radius_graph_fill_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, float r,
                              int32_t max_num_neighbors, int32_t loop, const int64_t* offsets,
                              int64_t* edge_index, int64_t n_edges, void* stream);
/* The same edge list (bit-identical) over the uniform cell list of gtb_dbscan_grid_f32 instead of the
 * all-pairs walk (the grid hash behind torch_cluster's radius search).  count builds the
 * cell list (cells r wide on the first min(d, 3) coordinates) in the workspace and counts; fill must be
 * given the SAME workspace, untouched in between, on the same stream.
 * workspace: gtb_radius_graph_grid_workspace_bytes(n) bytes, 256-byte aligned. */
size_t gtb_radius_graph_grid_workspace_bytes(int64_t n);
/* gtb_radius_pair_sum_f32 / gtb_radius_pair_sum_grad_f32 over the same cell list: same edges, same fp32
 * terms (sums in another order).  Each call builds the cell list in the workspace (same size as above). */
int gtb_radius_pair_sum_grid_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, const int64_t* pid,
                                 const uint8_t* src_flag, const float* beta, float q_min, float r, float p,
                                 float eps, int32_t max_num_neighbors, int32_t mode, double* out, void* workspace,
                                 size_t workspace_bytes, void* stream);
int gtb_radius_pair_sum_grad_grid_f32(const float* x, int32_t d, int64_t n, const int64_t* batch,
                                      const int64_t* pid, const uint8_t* src_flag, const float* beta, float q_min,
                                      float r, float p, float eps, int32_t max_num_neighbors, int32_t mode,
                                      const float* coef, float* gx, float* gq, void* workspace,
                                      size_t workspace_bytes, void* stream);
int gtb_radius_graph_grid_count_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, float r,
                                    int32_t max_num_neighbors, int32_t loop, int32_t* counts, void* workspace,
                                    size_t workspace_bytes, void* stream);
int gtb_radius_graph_grid_fill_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, float r,
                                   int32_t max_num_neighbors, int32_t loop, const int64_t* offsets,
                                   int64_t* edge_index, int64_t n_edges, void* workspace, size_t workspace_bytes,
                                   void* stream);

/* Gradients of the two sums above (what torch autograd derives for the reference's
 * norm / pow / relu chain over the radius-graph and true edges, metric_learning.py:14-54, oc.py:46-69;
 * d dist / d x = 0 at dist = 0).  coef: device float, the upstream gradient of the sum (already divided
 * by the loss normaliser).  gx [n, d] and gq [n] (mode 1: d / d charge, q = atanh(beta)^2 + q_min)
 * must be zero-filled (pair_sum) / are added onto (dist_pow); both kernels use atomics. */
int gtb_radius_pair_sum_grad_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, const int64_t* pid,
                                 const uint8_t* src_flag, const float* beta, float q_min, float r, float p,
                                 float eps, int32_t max_num_neighbors, int32_t mode, const float* coef,
                                 float* gx, float* gq, void* stream);
int gtb_edge_dist_pow_grad_f32(const float* x, int32_t d, const int64_t* edges, int64_t n_edges,
                               const uint8_t* src_flag, float p, const float* coef, float* gx, void* stream);

/* ----------------------------------------------------------------------- DBSCAN
 * One DBSCAN clustering of x [n, d] fp32 (d <= 16), replacing sklearn's radius_neighbors +
 * dbscan_inner as driven by DBSCANFastRescan.cluster (postprocessing/fastrescanner.py:40-66) inside the
 * hyper-parameter scan (postprocessing/dbscanscanner.py:146-187).  Neighbourhood dist <= eps
 * (float64, self included), core = at least min_pts neighbours.
 *   core   uint8 [n]: core-sample flags
 *   parent int32 [n]: scratch (union-find forest of the core samples, rooted at the lowest index)
 *   root   int32 [n]: lowest core index of the point's cluster; border points: the smallest adjacent
 *                     root; noise: -1.  sklearn's label = rank of `root` among the distinct roots. */
int gtb_dbscan_f32(const float* x, int32_t d, int64_t n, double eps, int32_t min_pts,
                   uint8_t* core, int32_t* parent, int32_t* root, void* stream);
/* The same clustering over a uniform cell list on the first min(d, 3) coordinates (cells at least eps wide:
 * the neighbour search SURVEY 2a K6 names) instead of the brute-force candidate walk; identical outputs.
 * workspace: gtb_dbscan_grid_workspace_bytes(n) bytes, 256-byte aligned. */
size_t gtb_dbscan_grid_workspace_bytes(int64_t n);
int gtb_dbscan_grid_f32(const float* x, int32_t d, int64_t n, double eps, int32_t min_pts,
                        uint8_t* core, int32_t* parent, int32_t* root,
                        void* workspace, size_t workspace_bytes, void* stream);

/* inv_norm[r] = 1 / max(||cat_s src_s[r]||_2, eps): torch.nn.functional.normalize(x, p=2, dim=1,
 * eps) as used by ResFCNN.forward (mlp.py:115-116); feed the result as row_scale. */
int gtb_rows_inv_l2norm_f32(const gtb_src_t* srcs, int32_t n_srcs, int64_t n_rows, float eps,
                            float* inv_norm, void* stream);

/* --------------------------------------------------------------------- row ops
 * Row gather / scatter used by the halo exchange of the node-partitioned
 * multi-GPU path (no reference equivalent: the reference is single-process). */
int gtb_rows_gather_f32(const float* src, int32_t src_ld, const int32_t* index, int64_t n_rows,
                        int32_t width, float* dst, int32_t dst_ld, void* stream);
int gtb_rows_scatter_f32(const float* src, int32_t src_ld, const int32_t* index, int64_t n_rows,
                         int32_t width, float* dst, int32_t dst_ld, void* stream);
/* dst[r] += src[index[r]]: the gradient of the per-destination aggregate (SumAggregation,
 * interaction_network.py:22,36) gathered back onto the edges and added to their own gradient. */
int gtb_rows_gather_add_f32(const float* src, int32_t src_ld, const int32_t* index, int64_t n_rows,
                            int32_t width, float* dst, int32_t dst_ld, void* stream);

/* ------------------------------------------------------------- backward blocks
 * out[Ka, Nb] += sum_r act(A[ia(r), 0:Ka])^T B[r, 0:Nb] and (colsum != NULL) colsum[Nb] += sum_r B[r]:
 * the weight and bias gradients of one Linear (torch autograd of models/mlp.py:59-62), with the
 * gather / ReLU-on-load of the forward source block (a_index NULL = identity rows).  Ka, Nb <= 64;
 * out / colsum are accumulated atomically (zero them first). */
int gtb_rows_atb_f32(const float* a, int32_t a_ld, const int32_t* a_index, int32_t a_relu, int32_t ka,
                     const float* b, int32_t b_ld, int32_t nb, int64_t n_rows,
                     float* out, int32_t out_ld, float* colsum, void* stream);
/* dst[index[r], 0:width] += src[r, 0:width]: the gradient of a row gather x[index]
 * (index_select inside MessagePassing.propagate, models/interaction_network.py:67). */
int gtb_rows_scatter_add_f32(const float* src, int32_t src_ld, const int32_t* index, int64_t n_rows,
                             int32_t width, float* dst, int32_t dst_ld, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GTB200_H */
