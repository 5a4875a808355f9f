// Graph plan: destination-sorted edge list (stable), CSR row pointers, and the plan of an
// edge sub-graph by stream compaction.  Integer work, bit-exact against
// torch.sort(stable=True) / bincount().cumsum() (oracle/in_oracle.py:plan).
#include <cub/cub.cuh>

#include "common.cuh"

namespace gtb {

__global__ void plan_keys_kernel(const int64_t* __restrict__ edge_index, int64_t n_nodes, int64_t n_edges,
                                 int32_t* __restrict__ keys, int32_t* __restrict__ vals,
                                 int32_t* __restrict__ status) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += stride) {
    const int64_t s = edge_index[e], t = edge_index[n_edges + e];
    if (s < 0 || s >= n_nodes || t < 0 || t >= n_nodes) atomicExch(status, 1);
    // out-of-range ids are reported through `status` and clamped: every kernel that walks the plan
    // stays inside its tables (the host raises the reference's IndexError, plan.py)
    keys[e] = (int32_t)max((int64_t)0, min(t, n_nodes - 1));
    vals[e] = (int32_t)e;
  }
}

// rowptr[n] = first sorted position whose destination is >= n; one thread per sorted edge
// fills the (possibly empty) run of nodes between its predecessor's destination and its own.
__global__ void plan_rowptr_kernel(const int32_t* __restrict__ dst_sorted, const int32_t* __restrict__ perm,
                                   const int64_t* __restrict__ edge_index, int64_t n_nodes, int64_t n_edges,
                                   int32_t* __restrict__ rowptr, int32_t* __restrict__ src_sorted) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_edges; i += stride) {
    const int64_t prev = (i == 0) ? -1 : dst_sorted[i - 1];
    const int64_t cur = (i == n_edges) ? n_nodes : min((int64_t)dst_sorted[i], n_nodes);
    for (int64_t n = max(prev, (int64_t)-1) + 1; n <= cur; ++n)
      if (n >= 0 && n <= n_nodes) rowptr[n] = (int32_t)i;
    if (i < n_edges && src_sorted) src_sorted[i] = (int32_t)max((int64_t)0, min(edge_index[perm[i]], n_nodes - 1));
  }
}

static int key_bits(int64_t n_nodes) {
  int b = 1;
  while (b < 31 && ((int64_t)1 << b) < n_nodes) ++b;
  return b;
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

size_t plan_workspace_bytes(int64_t n_nodes, int64_t n_edges) {
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n_edges, 0, key_bits(n_nodes));
  return align256(cub_bytes) + 2 * align256((size_t)n_edges * 4) + 256;
}

int plan_build(const int64_t* edge_index, int64_t n_nodes, int64_t n_edges, int32_t* perm, int32_t* rowptr,
               int32_t* src_sorted, int32_t* dst_sorted, int32_t* status, void* ws, size_t ws_bytes,
               cudaStream_t st) {
  GTB_REQUIRE(n_nodes >= 0 && n_edges >= 0 && n_nodes < (1ll << 31) - 1 && n_edges < (1ll << 31) - 1,
              GTB_ERR_BAD_ARG, "gtb_plan_build: n_nodes / n_edges out of the int32 range");
  GTB_REQUIRE(ws_bytes >= plan_workspace_bytes(n_nodes, n_edges), GTB_ERR_WORKSPACE,
              "gtb_plan_build: workspace too small");
  int rc = check_cuda(cudaMemsetAsync(status, 0, 4, st), "memset status");
  if (rc) return rc;
  const int threads = 256;
  const int blocks = (int)imin64((n_edges + threads) / threads, (int64_t)kNumSMs * 16);
  if (n_edges == 0) {
    plan_rowptr_kernel<<<1, threads, 0, st>>>(dst_sorted, perm, edge_index, n_nodes, 0, rowptr, src_sorted);
    GTB_CHECK_LAUNCH("plan_rowptr_kernel");
    return GTB_OK;
  }
  char* p = static_cast<char*>(ws);
  int32_t* keys = reinterpret_cast<int32_t*>(p);
  p += align256((size_t)n_edges * 4);
  int32_t* vals = reinterpret_cast<int32_t*>(p);
  p += align256((size_t)n_edges * 4);
  size_t cub_bytes = ws_bytes - 2 * align256((size_t)n_edges * 4);
  plan_keys_kernel<<<blocks, threads, 0, st>>>(edge_index, n_nodes, n_edges, keys, vals, status);
  GTB_CHECK_LAUNCH("plan_keys_kernel");
  // LSD radix sort is stable: equal destinations keep their original edge order, so per-destination
  // sums taken in sorted order reproduce the CPU scatter_add_ order.
  rc = check_cuda(cub::DeviceRadixSort::SortPairs(p, cub_bytes, keys, dst_sorted, vals, perm, (int)n_edges, 0,
                                                  key_bits(n_nodes), st),
                  "cub::DeviceRadixSort::SortPairs");
  if (rc) return rc;
  plan_rowptr_kernel<<<blocks, threads, 0, st>>>(dst_sorted, perm, edge_index, n_nodes, n_edges, rowptr,
                                                 src_sorted);
  GTB_CHECK_LAUNCH("plan_rowptr_kernel");
  return GTB_OK;
}

// ------------------------------------------------------------------ edge sub-graph plan
__global__ void filter_flags_kernel(const uint8_t* __restrict__ keep, const int32_t* __restrict__ perm,
                                    int64_t n_edges, int32_t* __restrict__ flag_sorted,
                                    int32_t* __restrict__ flag_orig) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_edges; i += stride) {
    flag_sorted[i] = keep[perm[i]] ? 1 : 0;
    flag_orig[i] = keep[i] ? 1 : 0;
  }
}

__global__ void filter_compact_kernel(const uint8_t* __restrict__ keep, const int32_t* __restrict__ perm,
                                      const int32_t* __restrict__ src_sorted, const int32_t* __restrict__ dst_sorted,
                                      const int32_t* __restrict__ pos_sorted /*exclusive scan*/,
                                      const int32_t* __restrict__ pos_orig /*exclusive scan*/, int64_t n_edges,
                                      int32_t* __restrict__ new_id, int32_t* __restrict__ kept_ids,
                                      int32_t* __restrict__ perm_out,
                                      int32_t* __restrict__ src_out, int32_t* __restrict__ dst_out,
                                      int32_t* __restrict__ n_kept) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_edges; i += stride) {
    new_id[i] = keep[i] ? pos_orig[i] : -1;
    if (keep[i]) kept_ids[pos_orig[i]] = (int32_t)i;
    const int32_t e = perm[i];
    if (keep[e]) {
      const int32_t j = pos_sorted[i];
      perm_out[j] = pos_orig[e];  // id of the edge inside the compacted (original-order) sub-graph
      src_out[j] = src_sorted[i];
      dst_out[j] = dst_sorted[i];
    }
    if (i == n_edges - 1) *n_kept = pos_orig[i] + (keep[i] ? 1 : 0);
  }
}

__global__ void rowptr_from_sorted_kernel(const int32_t* __restrict__ dst_sorted, const int32_t* __restrict__ n_kept_p,
                                          int64_t n_nodes, int64_t n_edges_max, int32_t* __restrict__ rowptr) {
  const int64_t n_edges = n_kept_p ? *n_kept_p : n_edges_max;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_edges; i += stride) {
    const int64_t prev = (i == 0) ? -1 : dst_sorted[i - 1];
    const int64_t cur = (i == n_edges) ? n_nodes : dst_sorted[i];
    for (int64_t n = prev + 1; n <= cur; ++n) rowptr[n] = (int32_t)i;
  }
}

size_t plan_filter_workspace_bytes(int64_t n_nodes, int64_t n_edges) {
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n_edges);
  return align256(cub_bytes) + 4 * align256((size_t)n_edges * 4) + 256;
}

int plan_filter(const uint8_t* keep, int64_t n_nodes, int64_t n_edges, const int32_t* perm,
                const int32_t* src_sorted, const int32_t* dst_sorted, int32_t* new_id, int32_t* kept_ids,
                int32_t* perm_out, int32_t* rowptr_out, int32_t* src_out, int32_t* dst_out, int32_t* n_kept_out, void* ws,
                size_t ws_bytes, cudaStream_t st) {
  GTB_REQUIRE(ws_bytes >= plan_filter_workspace_bytes(n_nodes, n_edges), GTB_ERR_WORKSPACE,
              "gtb_plan_filter: workspace too small");
  const int threads = 256;
  if (n_edges == 0) {
    int rc = check_cuda(cudaMemsetAsync(n_kept_out, 0, 4, st), "memset n_kept");
    if (rc) return rc;
    rowptr_from_sorted_kernel<<<1, threads, 0, st>>>(dst_out, nullptr, n_nodes, 0, rowptr_out);
    GTB_CHECK_LAUNCH("rowptr_from_sorted_kernel");
    return GTB_OK;
  }
  const int blocks = (int)imin64((n_edges + threads) / threads, (int64_t)kNumSMs * 16);
  char* p = static_cast<char*>(ws);
  const size_t seg = align256((size_t)n_edges * 4);
  int32_t* flag_sorted = reinterpret_cast<int32_t*>(p);
  int32_t* flag_orig = reinterpret_cast<int32_t*>(p + seg);
  int32_t* pos_sorted = reinterpret_cast<int32_t*>(p + 2 * seg);
  int32_t* pos_orig = reinterpret_cast<int32_t*>(p + 3 * seg);
  void* cub_ws = p + 4 * seg;
  size_t cub_bytes = ws_bytes - 4 * seg;
  filter_flags_kernel<<<blocks, threads, 0, st>>>(keep, perm, n_edges, flag_sorted, flag_orig);
  GTB_CHECK_LAUNCH("filter_flags_kernel");
  int rc = check_cuda(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, flag_sorted, pos_sorted, (int)n_edges, st),
                      "cub::DeviceScan::ExclusiveSum");
  if (rc) return rc;
  rc = check_cuda(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, flag_orig, pos_orig, (int)n_edges, st),
                  "cub::DeviceScan::ExclusiveSum");
  if (rc) return rc;
  filter_compact_kernel<<<blocks, threads, 0, st>>>(keep, perm, src_sorted, dst_sorted, pos_sorted, pos_orig,
                                                    n_edges, new_id, kept_ids, perm_out, src_out, dst_out, n_kept_out);
  GTB_CHECK_LAUNCH("filter_compact_kernel");
  rowptr_from_sorted_kernel<<<blocks, threads, 0, st>>>(dst_out, n_kept_out, n_nodes, n_edges, rowptr_out);
  GTB_CHECK_LAUNCH("rowptr_from_sorted_kernel");
  return GTB_OK;
}

// ----------------------------------------------------------------------------- orphan pruning
// Nodes without any edge are dropped and the others relabelled in increasing order (reference
// models/track_condensation_networks.py:254-259: unique endpoints of the surviving edges).  The relabelling is
// monotone, so the destination-sorted order of the plan survives: the plan of the pruned graph is this plan with
// relabelled endpoints and a compacted rowptr -- no second sort.
__global__ void prune_mark_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int64_t n_edges,
                                  int32_t* __restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_edges; i += stride) {
    flag[src[i]] = 1;  // same value from every writer
    flag[dst[i]] = 1;
  }
}

__global__ void prune_nodes_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ pos, const int32_t* __restrict__ rowptr,
                                   int64_t n_nodes, int32_t* __restrict__ new_id, int32_t* __restrict__ node_ids,
                                   int32_t* __restrict__ rowptr_out, int32_t* __restrict__ n_kept) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_nodes; v += stride) {
    const bool keep = flag[v] != 0;
    new_id[v] = keep ? pos[v] : -1;
    if (keep) {
      node_ids[pos[v]] = (int32_t)v;
      rowptr_out[pos[v]] = rowptr[v];  // a dropped node has no incoming edge: the segments in between are empty
    }
    if (v == n_nodes - 1) {
      const int32_t k = pos[v] + (keep ? 1 : 0);
      *n_kept = k;
      rowptr_out[k] = rowptr[n_nodes];
    }
  }
}

__global__ void prune_relabel_kernel(const int32_t* __restrict__ new_id, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                     int64_t n_edges, int32_t* __restrict__ src_out, int32_t* __restrict__ dst_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_edges; i += stride) {
    src_out[i] = new_id[src[i]];
    dst_out[i] = new_id[dst[i]];
  }
}

size_t plan_prune_workspace_bytes(int64_t n_nodes) {
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n_nodes);
  return align256(cub_bytes) + 2 * align256((size_t)n_nodes * 4) + 256;
}

int plan_prune(int64_t n_nodes, int64_t n_edges, const int32_t* rowptr, const int32_t* src_sorted, const int32_t* dst_sorted,
               int32_t* new_id, int32_t* node_ids, int32_t* rowptr_out, int32_t* src_out, int32_t* dst_out, int32_t* n_kept_out,
               void* ws, size_t ws_bytes, cudaStream_t st) {
  GTB_REQUIRE(n_nodes >= 0 && n_edges >= 0 && rowptr && new_id && node_ids && rowptr_out && n_kept_out && ws, GTB_ERR_BAD_ARG,
              "gtb_plan_prune_orphans: bad arguments");
  GTB_REQUIRE(ws_bytes >= plan_prune_workspace_bytes(n_nodes), GTB_ERR_WORKSPACE, "gtb_plan_prune_orphans: workspace too small");
  if (n_nodes == 0) return check_cuda(cudaMemsetAsync(n_kept_out, 0, 4, st), "memset n_kept");
  const int threads = 256;
  char* p = static_cast<char*>(ws);
  const size_t seg = align256((size_t)n_nodes * 4);
  int32_t* flag = reinterpret_cast<int32_t*>(p);
  int32_t* pos = reinterpret_cast<int32_t*>(p + seg);
  void* cub_ws = p + 2 * seg;
  size_t cub_bytes = ws_bytes - 2 * seg;
  int rc = check_cuda(cudaMemsetAsync(flag, 0, (size_t)n_nodes * 4, st), "memset flags");
  if (rc) return rc;
  const int eb = (int)imin64((n_edges + threads) / threads, (int64_t)kNumSMs * 16);
  const int nb = (int)imin64((n_nodes + threads) / threads, (int64_t)kNumSMs * 16);
  if (n_edges > 0) {
    prune_mark_kernel<<<eb, threads, 0, st>>>(src_sorted, dst_sorted, n_edges, flag);
    GTB_CHECK_LAUNCH("prune_mark_kernel");
  }
  rc = check_cuda(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, flag, pos, (int)n_nodes, st), "cub::DeviceScan::ExclusiveSum");
  if (rc) return rc;
  prune_nodes_kernel<<<nb, threads, 0, st>>>(flag, pos, rowptr, n_nodes, new_id, node_ids, rowptr_out, n_kept_out);
  GTB_CHECK_LAUNCH("prune_nodes_kernel");
  if (n_edges > 0) {
    prune_relabel_kernel<<<eb, threads, 0, st>>>(new_id, src_sorted, dst_sorted, n_edges, src_out, dst_out);
    GTB_CHECK_LAUNCH("prune_relabel_kernel");
  }
  return GTB_OK;
}

// ----------------------------------------------------------------------------- row ops
__global__ void rows_gather_kernel(const float* __restrict__ src, int src_ld, const int32_t* __restrict__ index,
                                   int64_t n_rows, int width, float* __restrict__ dst, int dst_ld, bool scatter) {
  const int64_t total = n_rows * width;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / width;
    const int c = (int)(i - r * width);
    if (scatter) dst[(size_t)index[r] * dst_ld + c] = src[(size_t)r * src_ld + c];
    else         dst[(size_t)r * dst_ld + c] = src[(size_t)index[r] * src_ld + c];
  }
}

// 16-byte pieces: thread = one float4 of a row.  MODE 0: dst[r] = src[index[r]], 1: dst[index[r]] = src[r],
// 2: dst[r] += src[index[r]] (the gathered per-node gradient added onto the per-edge one in the backward pass)
template <int MODE>
__global__ void rows_move_v4_kernel(const float* __restrict__ src, int src_ld, const int32_t* __restrict__ index,
                                    int64_t n_rows, int w4, float* __restrict__ dst, int dst_ld) {
  const int64_t total = n_rows * w4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / w4;
    const int c = (int)(i - r * w4) << 2;
    const int64_t ir = __ldg(index + r);
    if (MODE == 1) {
      *reinterpret_cast<float4*>(dst + (size_t)ir * dst_ld + c) = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * src_ld + c));
    } else {
      float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)ir * src_ld + c));
      float4* d = reinterpret_cast<float4*>(dst + (size_t)r * dst_ld + c);
      if (MODE == 2) {
        const float4 o = *d;
        v = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w);
      }
      *d = v;
    }
  }
}

__global__ void rows_gather_add_kernel(const float* __restrict__ src, int src_ld, const int32_t* __restrict__ index,
                                       int64_t n_rows, int width, float* __restrict__ dst, int dst_ld) {
  const int64_t total = n_rows * width;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / width;
    const int c = (int)(i - r * width);
    dst[(size_t)r * dst_ld + c] += src[(size_t)index[r] * src_ld + c];
  }
}

static bool rows_v4_ok(const float* src, int src_ld, int width, const float* dst, int dst_ld) {
  return (width & 3) == 0 && (src_ld & 3) == 0 && (dst_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
}

int rows_gather_add(const float* src, int src_ld, const int32_t* index, int64_t n_rows, int width, float* dst, int dst_ld,
                    cudaStream_t st) {
  GTB_REQUIRE(src && index && dst && width >= 1, GTB_ERR_BAD_ARG, "gtb_rows_gather_add_f32: bad arguments");
  if (n_rows == 0) return GTB_OK;
  if (rows_v4_ok(src, src_ld, width, dst, dst_ld)) {
    const int64_t total = n_rows * (width >> 2);
    const int blocks = (int)imin64((total + 255) / 256, (int64_t)kNumSMs * 16);
    rows_move_v4_kernel<2><<<blocks, 256, 0, st>>>(src, src_ld, index, n_rows, width >> 2, dst, dst_ld);
  } else {
    const int64_t total = n_rows * width;
    const int blocks = (int)imin64((total + 255) / 256, (int64_t)kNumSMs * 32);
    rows_gather_add_kernel<<<blocks, 256, 0, st>>>(src, src_ld, index, n_rows, width, dst, dst_ld);
  }
  GTB_CHECK_LAUNCH("rows_gather_add_kernel");
  return GTB_OK;
}

int rows_move(const float* src, int src_ld, const int32_t* index, int64_t n_rows, int width, float* dst,
              int dst_ld, bool scatter, cudaStream_t st) {
  if (n_rows == 0 || width == 0) return GTB_OK;
  if (rows_v4_ok(src, src_ld, width, dst, dst_ld)) {
    const int64_t total4 = n_rows * (width >> 2);
    const int blocks4 = (int)imin64((total4 + 255) / 256, (int64_t)kNumSMs * 16);
    if (scatter) rows_move_v4_kernel<1><<<blocks4, 256, 0, st>>>(src, src_ld, index, n_rows, width >> 2, dst, dst_ld);
    else         rows_move_v4_kernel<0><<<blocks4, 256, 0, st>>>(src, src_ld, index, n_rows, width >> 2, dst, dst_ld);
    GTB_CHECK_LAUNCH("rows_move_v4_kernel");
    return GTB_OK;
  }
  const int64_t total = n_rows * width;
  const int blocks = (int)imin64((total + 255) / 256, (int64_t)kNumSMs * 32);
  rows_gather_kernel<<<blocks, 256, 0, st>>>(src, src_ld, index, n_rows, width, dst, dst_ld, scatter);
  GTB_CHECK_LAUNCH("rows_gather_kernel");
  return GTB_OK;
}

}  // namespace gtb

namespace gtb {

struct NormSrcs {
  gtb_src_t s[GTB_MAX_SRCS];
  int n;
};

// one warp per row: sum of squares over every column block, then 1 / max(norm, eps)
__global__ void rows_inv_l2norm_kernel(const __grid_constant__ NormSrcs S, int64_t n_rows, float eps,
                                       float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += n_warps) {
    float ss = 0.f;
    for (int s = 0; s < S.n; ++s) {
      const int64_t row = S.s[s].index ? S.s[s].index[r] : r;
      const float* p = S.s[s].ptr + (size_t)row * S.s[s].ld;
      for (int c = lane; c < S.s[s].width; c += 32) {
        float v = p[c];
        if (S.s[s].relu) v = fmaxf(v, 0.f);
        ss = fmaf(v, v, ss);
      }
    }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) inv_norm[r] = 1.f / fmaxf(sqrtf(ss), eps);
  }
}

int rows_inv_l2norm(const gtb_src_t* srcs, int n_srcs, int64_t n_rows, float eps, float* inv_norm, cudaStream_t st) {
  GTB_REQUIRE(srcs && n_srcs >= 1 && n_srcs <= GTB_MAX_SRCS && inv_norm, GTB_ERR_BAD_ARG,
              "gtb_rows_inv_l2norm_f32: bad arguments");
  if (n_rows == 0) return GTB_OK;
  NormSrcs S;
  memset(&S, 0, sizeof(S));
  S.n = n_srcs;
  for (int i = 0; i < n_srcs; ++i) S.s[i] = srcs[i];
  const int blocks = (int)imin64((n_rows + 7) / 8, (int64_t)kNumSMs * 16);
  rows_inv_l2norm_kernel<<<blocks, 256, 0, st>>>(S, n_rows, eps, inv_norm);
  GTB_CHECK_LAUNCH("rows_inv_l2norm_kernel");
  return GTB_OK;
}

}  // namespace gtb
