// extern "C" surface of libgtb200 (see include/gtb200.h).
#include <stdarg.h>

#include "common.cuh"

namespace gtb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// implemented in the other translation units
size_t plan_workspace_bytes(int64_t, int64_t);
int plan_build(const int64_t*, int64_t, int64_t, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, void*, size_t,
               cudaStream_t);
size_t plan_filter_workspace_bytes(int64_t, int64_t);
int plan_filter(const uint8_t*, int64_t, int64_t, const int32_t*, const int32_t*, const int32_t*, int32_t*,
                int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, void*, size_t, cudaStream_t);
size_t plan_prune_workspace_bytes(int64_t);
int plan_prune(int64_t, int64_t, const int32_t*, const int32_t*, const int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*,
               int32_t*, void*, size_t, cudaStream_t);
int rows_inv_l2norm(const gtb_src_t*, int, int64_t, float, float*, cudaStream_t);
int rows_move(const float*, int, const int32_t*, int64_t, int, float*, int, bool, cudaStream_t);
int rows_gather_add(const float*, int, const int32_t*, int64_t, int, float*, int, cudaStream_t);
int pack_ffma(int, const int32_t*, const float* const*, const float* const*, void*, cudaStream_t);
int fused_mlp_ffma(const gtb_mlp_desc_t&, cudaStream_t);
size_t tc_packed_bytes(int, const int32_t*, int, const int32_t*);
int pack_tc(int, const int32_t*, int, const int32_t*, const float* const*, const float* const*, void*, cudaStream_t);
int fused_mlp_tc(const gtb_mlp_desc_t&, cudaStream_t);
int in_edge_ws(const gtb_mlp_desc_t&, cudaStream_t, bool*);
bool in_edge_ws_takes(const gtb_mlp_desc_t&);
bool ec_head_ws_takes(const gtb_mlp_desc_t&);
int ew_fault_flag(int*);
int ew_profile(int, long long*);
int ew_pack_bf16(const float* const*, const float* const*, void*, cudaStream_t);
int in_edge_ws_bf16(const void*, int32_t, const int32_t*, int32_t, const void*, int32_t, const void*, int32_t, int64_t,
                    const int32_t*, const int32_t*, const void*, void*, int32_t, const int32_t*, float*, int32_t, cudaStream_t);
int tc_timeout_flag(int*);
int in_node_ws(const float*, int32_t, int32_t, float*, int32_t, int32_t, int64_t, const void*, float, float, const float*, int32_t,
               float*, int32_t, const void*, const void*, int32_t, float*, int32_t, float*, int32_t, cudaStream_t);
int nw_fault_flag(int*);
int edge_encoder_ws(const float*, int32_t, const int32_t*, int64_t, int64_t, const float*, const float*, const void*, int32_t, float*,
                    int32_t, cudaStream_t);
int en_fault_flag(int*);
int at_fault_flag(int*);
int ec_head_ws(const gtb_mlp_desc_t&, cudaStream_t, bool*);
int hw_fault_flag(int*);
int oc_potentials_grad(const float*, const float*, int32_t, const int64_t*, const int32_t*, int64_t, const int32_t*, int32_t,
                       float, int64_t, const float*, float*, float*, float*, cudaStream_t);
int dbscan(const float*, int, int64_t, double, int, unsigned char*, int*, int*, cudaStream_t);
size_t dbscan_grid_workspace_bytes(int64_t);
int dbscan_grid(const float*, int, int64_t, double, int, unsigned char*, int*, int*, void*, size_t, cudaStream_t);
int radius_pair_sum(const float*, int, int64_t, const int64_t*, const int64_t*, const unsigned char*, const float*, float, float,
                    float, float, int, int, double*, cudaStream_t);
int edge_dist_pow_sum(const float*, int, const int64_t*, int64_t, const unsigned char*, float, double*, cudaStream_t);
int radius_graph_count(const float*, int, int64_t, const int64_t*, float, int, int, int32_t*, cudaStream_t);
int radius_graph_fill(const float*, int, int64_t, const int64_t*, float, int, int, const int64_t*, int64_t*, int64_t,
                      cudaStream_t);
int radius_pair_sum_grid(const float*, int, int64_t, const int64_t*, const int64_t*, const unsigned char*, const float*, float, float,
                         float, float, int, int, double*, const float*, float*, float*, void*, size_t, cudaStream_t);
int radius_graph_grid_count(const float*, int, int64_t, const int64_t*, float, int, int, int32_t*, void*, size_t, cudaStream_t);
int radius_graph_grid_fill(const float*, int, int64_t, const int64_t*, float, int, int, const int64_t*, int64_t*, int64_t, void*,
                           size_t, cudaStream_t);
int radius_pair_sum_grad(const float*, int, int64_t, const int64_t*, const int64_t*, const unsigned char*, const float*, float,
                         float, float, float, int, int, const float*, float*, float*, cudaStream_t);
int edge_dist_pow_grad(const float*, int, const int64_t*, int64_t, const unsigned char*, float, const float*, float*,
                       cudaStream_t);
int ec_loss_grad(const float*, const void*, int, int64_t, const int64_t*, const float*, float, int, float, float, float,
                 const float*, float*, cudaStream_t);
int rows_atb(const float*, int, const int32_t*, int, int, const float*, int, int, int64_t, float*, int, float*, cudaStream_t);
int rows_scatter_add(const float*, int, const int32_t*, int64_t, int, float*, int, cudaStream_t);
int tc_slots(int, const int32_t*, int, const int32_t*);
int tc_profile(int, long long*);
int ec_loss(const float*, const void*, int, int64_t, const int64_t*, const float*, float, int, float, float, float,
            double*, cudaStream_t);
size_t oc_workspace_bytes(int64_t);
int oc_prepare(const int64_t*, const uint8_t*, int64_t, int64_t*, int32_t*, int32_t*, void*, size_t, cudaStream_t);
int oc_alphas(const float*, const int32_t*, int64_t, float, int32_t, unsigned long long*, int32_t*, cudaStream_t);
int oc_potentials(const float*, const float*, int32_t, const int64_t*, const uint8_t*, const int32_t*, int64_t,
                  const int32_t*, int32_t, float, int64_t, double*, cudaStream_t);

static int validate_desc(const gtb_mlp_desc_t* d) {
  GTB_REQUIRE(d != nullptr, GTB_ERR_BAD_ARG, "gtb_fused_mlp_f32: null descriptor");
  GTB_REQUIRE(d->n_layers >= 1 && d->n_layers <= GTB_MAX_LAYERS, GTB_ERR_BAD_ARG,
              "gtb_fused_mlp_f32: n_layers=%d outside [1,%d]", d->n_layers, GTB_MAX_LAYERS);
  GTB_REQUIRE(d->n_srcs >= 1 && d->n_srcs <= GTB_MAX_SRCS, GTB_ERR_BAD_ARG,
              "gtb_fused_mlp_f32: n_srcs=%d outside [1,%d]", d->n_srcs, GTB_MAX_SRCS);
  GTB_REQUIRE(d->n_rows >= 0 && d->n_rows < (1ll << 31) - 256, GTB_ERR_BAD_ARG, "gtb_fused_mlp_f32: bad n_rows");
  int k = 0;
  for (int s = 0; s < d->n_srcs; ++s) {
    GTB_REQUIRE(d->srcs[s].ptr != nullptr && d->srcs[s].width >= 0 && d->srcs[s].ld >= d->srcs[s].width,
                GTB_ERR_BAD_ARG, "gtb_fused_mlp_f32: bad source block %d", s);
    if (d->srcs[s].flags & GTB_SRC_PROJECTED) {
      GTB_REQUIRE(d->n_layers >= 2 && d->srcs[s].width == d->dims[1] && !d->srcs[s].relu, GTB_ERR_BAD_ARG,
                  "gtb_fused_mlp_f32: pre-projected block %d must be %d wide, un-activated, in front of a hidden layer",
                  s, d->dims[1]);
    } else {
      k += d->srcs[s].width;
    }
  }
  // the reference asserts the feature widths at the same place (utils/asserts.py:4-7)
  GTB_REQUIRE(k == d->dims[0], GTB_ERR_BAD_ARG, "gtb_fused_mlp_f32: source widths sum to %d, first Linear expects %d",
              k, d->dims[0]);
  for (int l = 0; l <= d->n_layers; ++l)
    GTB_REQUIRE(d->dims[l] >= 1, GTB_ERR_BAD_ARG, "gtb_fused_mlp_f32: dims[%d]=%d", l, d->dims[l]);
  GTB_REQUIRE(d->packed != nullptr, GTB_ERR_BAD_ARG, "gtb_fused_mlp_f32: packed weights missing");
  GTB_REQUIRE(d->out != nullptr || d->aggr != nullptr, GTB_ERR_BAD_ARG, "gtb_fused_mlp_f32: no output requested");
  GTB_REQUIRE(d->aggr == nullptr || (d->seg_id != nullptr && d->rowptr != nullptr), GTB_ERR_BAD_ARG,
              "gtb_fused_mlp_f32: aggregate requested without seg_id / rowptr");
  GTB_REQUIRE(d->gate == nullptr || (d->gate_ld >= d->dims[d->n_layers] && d->aggr == nullptr), GTB_ERR_BAD_ARG,
              "gtb_fused_mlp_f32: bad gate (row stride below the output width, or combined with an aggregate)");
  return GTB_OK;
}

}  // namespace gtb

using namespace gtb;

extern "C" {

int gtb_version(void) { return 100; }
const char* gtb_last_error(void) { return g_err; }

int gtb_debug_tc_timeout(int* flag) {
  int a = 0, b = 0, c = 0;
  int rc = tc_timeout_flag(&a);
  if (rc == GTB_OK) rc = ew_fault_flag(&b);
  if (rc == GTB_OK) rc = nw_fault_flag(&c);
  int d = 0;
  if (rc == GTB_OK) rc = en_fault_flag(&d);
  int e = 0, f = 0;
  if (rc == GTB_OK) rc = at_fault_flag(&e);
  if (rc == GTB_OK) rc = hw_fault_flag(&f);
  *flag = a ? a : (b ? 16 + b : (c ? 32 + c : (d ? 48 + d : (e ? 64 + e : (f ? 80 + f : 0)))));
  return rc;
}
int gtb_debug_tc_profile(int enable, long long* out32) {
  // bit 0: the generic tcgen05 tiles, bit 1: the warp-specialised edge kernel (reads: enable == 2 selects it)
  if (out32 == nullptr) {
    const int rc = tc_profile(enable & 1, nullptr);
    return rc != GTB_OK ? rc : ew_profile((enable >> 1) & 1, nullptr);
  }
  return enable == 2 ? ew_profile(0, out32) : tc_profile(enable, out32);
}

size_t gtb_in_edge_bf16_packed_bytes(void) { return 99072; }

int gtb_in_edge_bf16_pack(const float* const* weights, const float* const* biases, void* packed, void* stream) {
  return ew_pack_bf16(weights, biases, packed, static_cast<cudaStream_t>(stream));
}

int gtb_in_edge_forward_bf16(const void* e_in, int32_t e_ld, const int32_t* e_index, int32_t relu_e, const void* p_i,
                             int32_t pi_ld, const void* p_j, int32_t pj_ld, int64_t n_edges, const int32_t* src_sorted,
                             const int32_t* dst_sorted, const void* packed, void* e_out, int32_t eo_ld,
                             const int32_t* out_index, float* aggr, int32_t aggr_ld, void* stream) {
  return in_edge_ws_bf16(e_in, e_ld, e_index, relu_e, p_i, pi_ld, p_j, pj_ld, n_edges, src_sorted, dst_sorted, packed, e_out,
                         eo_ld, out_index, aggr, aggr_ld, static_cast<cudaStream_t>(stream));
}

int gtb_in_node_fused_f32(const float* x, int32_t x_ld, int32_t relu_x, float* aggr, int32_t aggr_ld, int32_t zero_aggr,
                          int64_t n_nodes, const void* packed_obj, float res_a, float res_b, const float* res, int32_t res_ld,
                          float* x_out, int32_t xo_ld, const void* packed_pa, const void* packed_pb, int32_t proj_relu,
                          float* p_a, int32_t pa_ld, float* p_b, int32_t pb_ld, void* stream) {
  return in_node_ws(x, x_ld, relu_x, aggr, aggr_ld, zero_aggr, n_nodes, packed_obj, res_a, res_b, res, res_ld, x_out, xo_ld,
                    packed_pa, packed_pb, proj_relu, p_a, pa_ld, p_b, pb_ld, static_cast<cudaStream_t>(stream));
}

int gtb_edge_encoder_f32(const float* x, int32_t x_ld, const int32_t* index, int64_t n_rows, int64_t x_rows, const float* w0,
                         const float* b0, const void* packed_w1, int32_t final_relu, float* out, int32_t out_ld, void* stream) {
  return edge_encoder_ws(x, x_ld, index, n_rows, x_rows, w0, b0, packed_w1, final_relu, out, out_ld, static_cast<cudaStream_t>(stream));
}

size_t gtb_dbscan_grid_workspace_bytes(int64_t n) { return dbscan_grid_workspace_bytes(n); }

int gtb_dbscan_grid_f32(const float* x, int32_t d, int64_t n, double eps, int32_t min_pts, uint8_t* core, int32_t* parent,
                        int32_t* root, void* workspace, size_t workspace_bytes, void* stream) {
  return dbscan_grid(x, d, n, eps, min_pts, core, parent, root, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int gtb_arch_ok(int device) {
  cudaDeviceProp p;
  int rc = check_cuda(cudaGetDeviceProperties(&p, device), "cudaGetDeviceProperties");
  if (rc) return rc;
  GTB_REQUIRE(p.major == 10, GTB_ERR_ARCH, "gtb200 needs a compute-capability 10.x GPU (B200, sm_100a); device %d is %d.%d",
              device, p.major, p.minor);
  return GTB_OK;
}

size_t gtb_plan_workspace_bytes(int64_t n_nodes, int64_t n_edges) { return plan_workspace_bytes(n_nodes, n_edges); }

int gtb_plan_build(const int64_t* edge_index, int64_t n_nodes, int64_t n_edges, int32_t* perm, int32_t* rowptr,
                   int32_t* src_sorted, int32_t* dst_sorted, int32_t* status, void* workspace,
                   size_t workspace_bytes, void* stream) {
  return plan_build(edge_index, n_nodes, n_edges, perm, rowptr, src_sorted, dst_sorted, status, workspace,
                    workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t gtb_plan_filter_workspace_bytes(int64_t n_nodes, int64_t n_edges) {
  return plan_filter_workspace_bytes(n_nodes, n_edges);
}

int gtb_plan_filter(const uint8_t* keep, int64_t n_nodes, int64_t n_edges, const int32_t* perm,
                    const int32_t* src_sorted, const int32_t* dst_sorted, int32_t* new_id, int32_t* kept_ids,
                    int32_t* perm_out, int32_t* rowptr_out, int32_t* src_sorted_out, int32_t* dst_sorted_out, int32_t* n_kept_out,
                    void* workspace, size_t workspace_bytes, void* stream) {
  return plan_filter(keep, n_nodes, n_edges, perm, src_sorted, dst_sorted, new_id, kept_ids, perm_out, rowptr_out,
                     src_sorted_out, dst_sorted_out, n_kept_out, workspace, workspace_bytes,
                     static_cast<cudaStream_t>(stream));
}

size_t gtb_mlp_packed_bytes(int n_layers, const int32_t* dims, int n_blocks, const int32_t* block_widths, int impl) {
  if (n_layers < 1 || n_layers > GTB_MAX_LAYERS || dims == nullptr) return 0;
  if (impl == GTB_IMPL_TCGEN05) return tc_packed_bytes(n_layers, dims, n_blocks, block_widths);
  FfmaLayout L;
  if (!ffma_layout(n_layers, dims, &L)) return 0;
  return L.total_floats * sizeof(float);
}

int gtb_mlp_tc_slots(int n_layers, const int32_t* dims, int n_blocks, const int32_t* block_widths) {
  if (n_layers < 1 || n_layers > GTB_MAX_LAYERS || dims == nullptr) return 0;
  return tc_slots(n_layers, dims, n_blocks, block_widths);
}

int gtb_mlp_pack(int n_layers, const int32_t* dims, int n_blocks, const int32_t* block_widths,
                 const float* const* weights, const float* const* biases, int impl, void* packed, void* stream) {
  GTB_REQUIRE(n_layers >= 1 && n_layers <= GTB_MAX_LAYERS && dims && weights && packed, GTB_ERR_BAD_ARG,
              "gtb_mlp_pack: bad arguments");
  if (impl == GTB_IMPL_TCGEN05)
    return pack_tc(n_layers, dims, n_blocks, block_widths, weights, biases, packed, static_cast<cudaStream_t>(stream));
  GTB_REQUIRE(impl == GTB_IMPL_FFMA, GTB_ERR_BAD_ARG, "gtb_mlp_pack: impl must be GTB_IMPL_FFMA or GTB_IMPL_TCGEN05");
  return pack_ffma(n_layers, dims, weights, biases, packed, static_cast<cudaStream_t>(stream));
}

int gtb_fused_mlp_f32(const gtb_mlp_desc_t* desc, void* stream) {
  int rc = validate_desc(desc);
  if (rc) return rc;
  GTB_REQUIRE(desc->impl == GTB_IMPL_FFMA || desc->impl == GTB_IMPL_TCGEN05, GTB_ERR_BAD_ARG,
              "gtb_fused_mlp_f32: impl must name the layout the weights were packed for");
  if (desc->impl == GTB_IMPL_TCGEN05) {
    // the wide Interaction-Network edge shape has its own warp-specialised kernel (edge_ws.cu)
    bool handled = false;
    int rc = in_edge_ws(*desc, static_cast<cudaStream_t>(stream), &handled);
    if (rc != GTB_OK || handled) return rc;
    // ... and so has the wide W head of the edge classifier (head_ws.cu)
    rc = ec_head_ws(*desc, static_cast<cudaStream_t>(stream), &handled);
    if (rc != GTB_OK || handled) return rc;
    GTB_REQUIRE(desc->hidden0 == nullptr && desc->hidden1 == nullptr, GTB_ERR_UNSUPPORTED_DIM,
                "gtb_fused_mlp_f32: hidden0 / hidden1 requested for a launch the generic tiles run (see gtb_fused_mlp_saves_hidden)");
    return fused_mlp_tc(*desc, static_cast<cudaStream_t>(stream));
  }
  GTB_REQUIRE(desc->hidden0 == nullptr && desc->hidden1 == nullptr, GTB_ERR_UNSUPPORTED_DIM,
              "gtb_fused_mlp_f32: hidden0 / hidden1 requested for a launch the FFMA tiles run (see gtb_fused_mlp_saves_hidden)");
  return fused_mlp_ffma(*desc, static_cast<cudaStream_t>(stream));
}

int gtb_fused_mlp_saves_hidden(const gtb_mlp_desc_t* desc) {
  if (desc == nullptr || desc->impl != GTB_IMPL_TCGEN05 || validate_desc(desc) != GTB_OK) return 0;
  return (in_edge_ws_takes(*desc) || ec_head_ws_takes(*desc)) ? 1 : 0;
}

int gtb_in_edge_forward_f32(const float* x, int32_t x_ld, int32_t relu_x, const float* edge_attr, int32_t e_ld,
                            int32_t relu_e, int64_t n_nodes, int64_t n_edges, const int32_t* perm,
                            const int32_t* rowptr, const int32_t* src_sorted, const int32_t* dst_sorted,
                            int32_t node_dim, int32_t edge_dim, int32_t hidden, int32_t edge_outdim,
                            const void* packed_rel, int impl, float* e_tilde, int32_t eo_ld, float* aggr,
                            void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GTB_REQUIRE(aggr != nullptr, GTB_ERR_BAD_ARG, "gtb_in_edge_forward_f32: aggr is required");
  int rc = check_cuda(cudaMemsetAsync(aggr, 0, (size_t)n_nodes * edge_outdim * sizeof(float), st), "memset aggr");
  if (rc) return rc;
  if (n_edges == 0) return GTB_OK;
  gtb_mlp_desc_t d;
  memset(&d, 0, sizeof(d));
  d.n_rows = n_edges;
  d.n_srcs = 3;
  d.n_layers = 3;
  // message(): cat[x_i, x_j, edge_attr], x_i = target rows, x_j = source rows (interaction_network.py:75-86)
  d.srcs[0] = gtb_src_t{x, dst_sorted, node_dim, x_ld, relu_x, 0};
  d.srcs[1] = gtb_src_t{x, src_sorted, node_dim, x_ld, relu_x, 0};
  d.srcs[2] = gtb_src_t{edge_attr, perm, edge_dim, e_ld, relu_e, 0};
  d.dims[0] = 2 * node_dim + edge_dim;
  d.dims[1] = hidden;
  d.dims[2] = hidden;
  d.dims[3] = edge_outdim;
  d.packed = packed_rel;
  d.impl = impl;
  d.final_act = GTB_ACT_NONE;
  d.res_b = 1.f;
  d.out = e_tilde;
  d.out_index = perm;
  d.out_ld = eo_ld;
  d.aggr = aggr;
  d.aggr_ld = edge_outdim;
  d.seg_id = dst_sorted;
  d.rowptr = rowptr;
  return gtb_fused_mlp_f32(&d, stream);
}

int gtb_in_node_forward_f32(const float* x, int32_t x_ld, int32_t relu_x, const float* aggr, int64_t n_nodes,
                            int32_t node_dim, int32_t aggr_dim, int32_t hidden, int32_t node_outdim,
                            const void* packed_obj, int impl, float res_a, float res_b, const float* res,
                            int32_t res_ld, float* x_out, int32_t xo_ld, void* stream) {
  if (n_nodes == 0) return GTB_OK;
  gtb_mlp_desc_t d;
  memset(&d, 0, sizeof(d));
  d.n_rows = n_nodes;
  d.n_srcs = 2;
  d.n_layers = 3;
  // update(): cat[x, aggr_out] (interaction_network.py:92-103)
  d.srcs[0] = gtb_src_t{x, nullptr, node_dim, x_ld, relu_x, 0};
  d.srcs[1] = gtb_src_t{aggr, nullptr, aggr_dim, aggr_dim, 0, 0};
  d.dims[0] = node_dim + aggr_dim;
  d.dims[1] = hidden;
  d.dims[2] = hidden;
  d.dims[3] = node_outdim;
  d.packed = packed_obj;
  d.impl = impl;
  d.final_act = GTB_ACT_NONE;
  d.res = res;
  d.res_ld = res_ld;
  d.res_a = res_a;
  d.res_b = res_b;
  d.out = x_out;
  d.out_ld = xo_ld;
  return gtb_fused_mlp_f32(&d, stream);
}

int gtb_ec_loss_f32(const float* w, const void* y, int label_kind, int64_t n_edges, const int64_t* src,
                    const float* pt, float pt_thld, int mode, float alpha, float gamma, float pos_weight,
                    double* out, void* stream) {
  return ec_loss(w, y, label_kind, n_edges, src, pt, pt_thld, mode, alpha, gamma, pos_weight, out,
                 static_cast<cudaStream_t>(stream));
}

int gtb_ec_loss_grad_f32(const float* w, const void* y, int label_kind, int64_t n_edges, const int64_t* src,
                         const float* pt, float pt_thld, int mode, float alpha, float gamma, float pos_weight,
                         const float* scale, float* dw, void* stream) {
  return ec_loss_grad(w, y, label_kind, n_edges, src, pt, pt_thld, mode, alpha, gamma, pos_weight, scale, dw,
                      static_cast<cudaStream_t>(stream));
}

size_t gtb_oc_workspace_bytes(int64_t n_nodes) { return oc_workspace_bytes(n_nodes); }

int gtb_oc_prepare(const int64_t* object_id, const uint8_t* object_mask, int64_t n_nodes, int64_t* uniq,
                   int32_t* obj_slot, int32_t* n_uniq, void* workspace, size_t workspace_bytes, void* stream) {
  return oc_prepare(object_id, object_mask, n_nodes, uniq, obj_slot, n_uniq, workspace, workspace_bytes,
                    static_cast<cudaStream_t>(stream));
}

int gtb_oc_alphas(const float* beta, const int32_t* obj_slot, int64_t n_nodes, float q_min, int32_t k,
                  unsigned long long* packed_scratch, int32_t* alphas, void* stream) {
  return oc_alphas(beta, obj_slot, n_nodes, q_min, k, packed_scratch, alphas, static_cast<cudaStream_t>(stream));
}

int gtb_oc_potentials(const float* beta, const float* x, int32_t d, const int64_t* object_id,
                      const uint8_t* object_mask, const int32_t* obj_slot, int64_t n_nodes, const int32_t* alphas,
                      int32_t k, float q_min, int64_t noise_threshold, double* out, void* stream) {
  return oc_potentials(beta, x, d, object_id, object_mask, obj_slot, n_nodes, alphas, k, q_min, noise_threshold,
                       out, static_cast<cudaStream_t>(stream));
}

int gtb_radius_pair_sum_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, const int64_t* pid,
                            const uint8_t* src_flag, const float* beta, float q_min, float r, float p, float eps,
                            int32_t max_num_neighbors, int32_t mode, double* out, void* stream) {
  return radius_pair_sum(x, d, n, batch, pid, src_flag, beta, q_min, r, p, eps, max_num_neighbors, mode, out,
                         static_cast<cudaStream_t>(stream));
}

int gtb_radius_graph_count_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, float r, int32_t max_num_neighbors,
                               int32_t loop, int32_t* counts, void* stream) {
  return radius_graph_count(x, d, n, batch, r, max_num_neighbors, loop, counts, static_cast<cudaStream_t>(stream));
}

int gtb_radius_graph_fill_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, float r, int32_t max_num_neighbors,
                              int32_t loop, const int64_t* offsets, int64_t* edge_index, int64_t n_edges, void* stream) {
  return radius_graph_fill(x, d, n, batch, r, max_num_neighbors, loop, offsets, edge_index, n_edges,
                           static_cast<cudaStream_t>(stream));
}

size_t gtb_radius_graph_grid_workspace_bytes(int64_t n) { return dbscan_grid_workspace_bytes(n); }

int gtb_radius_pair_sum_grid_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, const int64_t* pid,
                                 const uint8_t* src_flag, const float* beta, float q_min, float r, float p, float eps,
                                 int32_t max_num_neighbors, int32_t mode, double* out, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  GTB_REQUIRE(out != nullptr, GTB_ERR_BAD_ARG, "gtb_radius_pair_sum_grid_f32: out is null");
  return radius_pair_sum_grid(x, d, n, batch, pid, src_flag, beta, q_min, r, p, eps, max_num_neighbors, mode, out, nullptr,
                              nullptr, nullptr, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int gtb_radius_pair_sum_grad_grid_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, const int64_t* pid,
                                      const uint8_t* src_flag, const float* beta, float q_min, float r, float p, float eps,
                                      int32_t max_num_neighbors, int32_t mode, const float* coef, float* gx, float* gq,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  GTB_REQUIRE(coef != nullptr, GTB_ERR_BAD_ARG, "gtb_radius_pair_sum_grad_grid_f32: coef is null");
  return radius_pair_sum_grid(x, d, n, batch, pid, src_flag, beta, q_min, r, p, eps, max_num_neighbors, mode, nullptr, coef, gx,
                              gq, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int gtb_radius_graph_grid_count_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, float r,
                                    int32_t max_num_neighbors, int32_t loop, int32_t* counts, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  return radius_graph_grid_count(x, d, n, batch, r, max_num_neighbors, loop, counts, workspace, workspace_bytes,
                                 static_cast<cudaStream_t>(stream));
}

int gtb_radius_graph_grid_fill_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, float r,
                                   int32_t max_num_neighbors, int32_t loop, const int64_t* offsets, int64_t* edge_index,
                                   int64_t n_edges, void* workspace, size_t workspace_bytes, void* stream) {
  return radius_graph_grid_fill(x, d, n, batch, r, max_num_neighbors, loop, offsets, edge_index, n_edges, workspace,
                                workspace_bytes, static_cast<cudaStream_t>(stream));
}

int gtb_radius_pair_sum_grad_f32(const float* x, int32_t d, int64_t n, const int64_t* batch, const int64_t* pid,
                                 const uint8_t* src_flag, const float* beta, float q_min, float r, float p, float eps,
                                 int32_t max_num_neighbors, int32_t mode, const float* coef, float* gx, float* gq,
                                 void* stream) {
  return radius_pair_sum_grad(x, d, n, batch, pid, src_flag, beta, q_min, r, p, eps, max_num_neighbors, mode, coef, gx, gq,
                              static_cast<cudaStream_t>(stream));
}

int gtb_edge_dist_pow_grad_f32(const float* x, int32_t d, const int64_t* edges, int64_t n_edges, const uint8_t* src_flag,
                               float p, const float* coef, float* gx, void* stream) {
  return edge_dist_pow_grad(x, d, edges, n_edges, src_flag, p, coef, gx, static_cast<cudaStream_t>(stream));
}

int gtb_edge_dist_pow_sum_f32(const float* x, int32_t d, const int64_t* edges, int64_t n_edges, const uint8_t* src_flag,
                              float p, double* out, void* stream) {
  return edge_dist_pow_sum(x, d, edges, n_edges, src_flag, p, out, static_cast<cudaStream_t>(stream));
}

int gtb_oc_potentials_grad(const float* beta, const float* x, int32_t d, const int64_t* object_id, const int32_t* obj_slot,
                           int64_t n_nodes, const int32_t* alphas, int32_t k, float q_min, int64_t noise_threshold,
                           const float* coef, float* gq, float* gbeta, float* gx, void* stream) {
  return oc_potentials_grad(beta, x, d, object_id, obj_slot, n_nodes, alphas, k, q_min, noise_threshold, coef, gq, gbeta, gx,
                            static_cast<cudaStream_t>(stream));
}

int gtb_dbscan_f32(const float* x, int32_t d, int64_t n, double eps, int32_t min_pts, uint8_t* core, int32_t* parent,
                   int32_t* root, void* stream) {
  return dbscan(x, d, n, eps, min_pts, core, parent, root, static_cast<cudaStream_t>(stream));
}

int gtb_rows_inv_l2norm_f32(const gtb_src_t* srcs, int32_t n_srcs, int64_t n_rows, float eps, float* inv_norm,
                            void* stream) {
  return rows_inv_l2norm(srcs, n_srcs, n_rows, eps, inv_norm, static_cast<cudaStream_t>(stream));
}

size_t gtb_plan_prune_workspace_bytes(int64_t n_nodes) { return plan_prune_workspace_bytes(n_nodes); }

int gtb_plan_prune_orphans(int64_t n_nodes, int64_t n_edges, const int32_t* rowptr, const int32_t* src_sorted,
                           const int32_t* dst_sorted, int32_t* new_id, int32_t* node_ids, int32_t* rowptr_out, int32_t* src_out,
                           int32_t* dst_out, int32_t* n_kept_out, void* workspace, size_t workspace_bytes, void* stream) {
  return plan_prune(n_nodes, n_edges, rowptr, src_sorted, dst_sorted, new_id, node_ids, rowptr_out, src_out, dst_out, n_kept_out,
                    workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int gtb_rows_gather_f32(const float* src, int32_t src_ld, const int32_t* index, int64_t n_rows, int32_t width,
                        float* dst, int32_t dst_ld, void* stream) {
  return rows_move(src, src_ld, index, n_rows, width, dst, dst_ld, false, static_cast<cudaStream_t>(stream));
}

int gtb_rows_gather_add_f32(const float* src, int32_t src_ld, const int32_t* index, int64_t n_rows, int32_t width,
                            float* dst, int32_t dst_ld, void* stream) {
  return rows_gather_add(src, src_ld, index, n_rows, width, dst, dst_ld, static_cast<cudaStream_t>(stream));
}

int gtb_rows_scatter_f32(const float* src, int32_t src_ld, const int32_t* index, int64_t n_rows, int32_t width,
                         float* dst, int32_t dst_ld, void* stream) {
  return rows_move(src, src_ld, index, n_rows, width, dst, dst_ld, true, static_cast<cudaStream_t>(stream));
}

int gtb_rows_atb_f32(const float* a, int32_t a_ld, const int32_t* a_index, int32_t a_relu, int32_t ka, const float* b,
                     int32_t b_ld, int32_t nb, int64_t n_rows, float* out, int32_t out_ld, float* colsum, void* stream) {
  return rows_atb(a, a_ld, a_index, a_relu, ka, b, b_ld, nb, n_rows, out, out_ld, colsum, static_cast<cudaStream_t>(stream));
}

int gtb_rows_scatter_add_f32(const float* src, int32_t src_ld, const int32_t* index, int64_t n_rows, int32_t width,
                             float* dst, int32_t dst_ld, void* stream) {
  return rows_scatter_add(src, src_ld, index, n_rows, width, dst, dst_ld, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
