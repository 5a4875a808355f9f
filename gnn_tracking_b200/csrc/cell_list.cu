// Uniform cell list over a low-dimensional latent space (the neighbour search SURVEY 2a K6 / 8f-2 / 8f-3 names), and
// the two fixed-radius searches built on it: the DBSCAN trial (same results as dbscan.cu -- labels identical to
// sklearn's) and the materialised radius graph (same edge list as radius.cu's all-pairs walk), without the N^2
// candidate walk.
//
// The points are binned on their first min(d, 3) coordinates into cells at least eps wide, so every
// neighbour (dist <= eps) of a point lies in the 3^k cells around its own; the full-dimensional distance
// tests are the ones of dbscan.cu (fp32 screen, float64 decision on the rim) and radius.cu (fp32, strict).
// Per build (DBSCAN: one trial of the hyper-parameter scan, postprocessing/dbscanscanner.py:146-187; eps changes
// from trial to trial, so the grid is rebuilt -- a 21-bit radix sort of the cell ids):
//   bounding box -> cell ids -> sort (cell id, point) -> cell_begin[] -> coordinates in cell order
// DBSCAN passes:
//   pass 0: neighbour counts -> core flags        pass 1: union-find of the core samples
//   pass 2: roots (core: own component, border: smallest adjacent root, noise: -1)
// Everything visible to the caller (core, parent, root) is indexed by the ORIGINAL point index, and the
// outcome does not depend on the order in which neighbours are visited (components are rooted at their
// lowest index, a border point takes the smallest adjacent root), so the numbering equals dbscan_inner's.
// Radius-graph passes: count (neighbours per centre, capped), fill (sorted insertion into the centre's segment
// of the edge list: ascending neighbours, the lowest indices kept under the cap -- radius.cu's order).
#include <cub/cub.cuh>

#include "common.cuh"

namespace gtb {

constexpr int DG_T = 128;
constexpr int DG_MAXD = 16;
constexpr int DG_CELL_BITS = 21;  // at most 2^21 cells

struct DgGrid {        // written by dg_setup_kernel, read by everything else
  float lo[3];         // lower corner of the box in the grid dimensions
  double inv_h[3];     // 1 / cell size
  int g[3];            // cells per grid dimension (unused dimensions: 1)
  int dim[3];          // coordinate index of each grid dimension (-1: unused)
  int n_cells;
};

__device__ __forceinline__ unsigned dg_f2o(float f) {  // order-preserving float -> unsigned
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dg_o2f(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void dg_bbox_kernel(const float* __restrict__ x, int d, int64_t n, int k, unsigned* __restrict__ mm /* [3 min | 3 max] */) {
  unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    for (int c = 0; c < k; ++c) {
      const unsigned o = dg_f2o(__ldg(x + i * d + c));
      lo[c] = min(lo[c], o);
      hi[c] = max(hi[c], o);
    }
  for (int c = 0; c < k; ++c) {
    for (int s = 16; s > 0; s >>= 1) {
      lo[c] = min(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], s));
      hi[c] = max(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], s));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(mm + c, lo[c]);
      atomicMax(mm + 3 + c, hi[c]);
    }
  }
}

// grid dimensions: the LAST grid dimension is always used (cells adjacent in it are adjacent in memory)
__global__ void dg_setup_kernel(const unsigned* __restrict__ mm, int k, double eps, DgGrid* __restrict__ grid) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int cap = k == 1 ? (1 << DG_CELL_BITS) : (k == 2 ? 1448 : 128);  // cap^k <= 2^21
  DgGrid gr;
  gr.n_cells = 1;
  for (int a = 0; a < 3; ++a) {
    const int c = a - (3 - k);  // coordinate index, or < 0 for an unused leading dimension
    gr.dim[a] = c;
    gr.lo[a] = 0.f;
    gr.inv_h[a] = 0.0;
    gr.g[a] = 1;
    if (c < 0) continue;
    const float lo = dg_o2f(mm[c]), hi = dg_o2f(mm[3 + c]);
    double range = (double)hi - (double)lo;
    if (!(range >= 0.0) || !isfinite(range)) range = 0.0;  // NaN / Inf coordinates: one cell (they match nothing anyway)
    double h = eps * (1.0 + 1e-9);
    if (!(h > 0.0)) h = 1.0;
    int g = range / h >= (double)cap ? cap : (int)(range / h) + 1;
    if (g < 1) g = 1;
    if (g == cap) h = range / (double)cap * (1.0 + 1e-9);  // wider cells than eps: still a valid cell list
    gr.lo[a] = lo;
    gr.inv_h[a] = 1.0 / h;
    gr.g[a] = g;
    gr.n_cells *= g;
  }
  *grid = gr;
}

__device__ __forceinline__ int dg_cell_coord(float v, const DgGrid& gr, int a) {
  int c = (int)floor(((double)v - (double)gr.lo[a]) * gr.inv_h[a]);
  return c < 0 ? 0 : (c >= gr.g[a] ? gr.g[a] - 1 : c);  // NaNs land in cell 0
}

__global__ void dg_keys_kernel(const float* __restrict__ x, int d, int64_t n, const DgGrid* __restrict__ grid,
                               int32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const DgGrid gr = *grid;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int cell = 0;
    for (int a = 0; a < 3; ++a) {
      const int c = gr.dim[a] < 0 ? 0 : dg_cell_coord(__ldg(x + i * d + gr.dim[a]), gr, a);
      cell = cell * gr.g[a] + c;
    }
    keys[i] = cell;
    vals[i] = (int32_t)i;
  }
}

// cell_begin[c] = first sorted position whose cell is >= c (c in [0, n_cells]); coordinates in cell order,
// zero-padded to D floats
template <int D>
__global__ void dg_layout_kernel(const float* __restrict__ x, int d, int64_t n, const DgGrid* __restrict__ grid,
                                 const int32_t* __restrict__ keys_sorted, const int32_t* __restrict__ idx_sorted,
                                 int32_t* __restrict__ cell_begin, float* __restrict__ xs) {
  const int n_cells = grid->n_cells;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p <= n; p += stride) {
    const int prev = p == 0 ? -1 : keys_sorted[p - 1];
    const int cur = p == n ? n_cells : keys_sorted[p];
    for (int c = prev + 1; c <= cur; ++c) cell_begin[c] = (int32_t)p;
    if (p < n) {
      const int64_t i = idx_sorted[p];
#pragma unroll
      for (int c = 0; c < D; ++c) xs[p * D + c] = c < d ? __ldg(x + i * d + c) : 0.f;
    }
  }
}

template <int D>
__device__ __forceinline__ bool dg_near(const float (&xi)[D], const float* __restrict__ xj, double eps2, float e2_hi, float e2_lo) {
  float v[D];
#pragma unroll
  for (int q = 0; q < D / 4; ++q) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(xj) + q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    const float u = xi[c] - v[c];
    s = fmaf(u, u, s);
  }
  if (s > e2_hi) return false;
  if (s < e2_lo) return true;
  double d2 = 0.0;  // on the rim of the eps-ball: decide in float64 like sklearn
#pragma unroll
  for (int c = 0; c < D; ++c) {
    const double u = (double)xi[c] - (double)v[c];
    d2 = fma(u, u, d2);
  }
  return d2 <= eps2;
}

__device__ __forceinline__ int dg_find(int* parent, int v) {
  while (true) {
    const int p = __ldcg(parent + v);
    if (p == v) return v;
    const int gp = __ldcg(parent + p);
    if (gp != p) __stcg(parent + v, gp);
    v = p;
  }
}
__device__ __forceinline__ void dg_unite(int* parent, int a, int b) {
  while (true) {
    a = dg_find(parent, a);
    b = dg_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    if (atomicCAS(parent + a, a, b) == a) return;  // the larger root goes under the smaller one
  }
}

// one thread per point, walked in cell order (neighbouring threads share their candidate ranges)
template <int D>
__global__ void __launch_bounds__(DG_T) dg_pass_kernel(const float* __restrict__ xs, int64_t n, const DgGrid* __restrict__ grid,
                                                       const int32_t* __restrict__ keys_sorted,
                                                       const int32_t* __restrict__ idx_sorted,
                                                       const int32_t* __restrict__ cell_begin, double eps2, int min_pts, int phase,
                                                       unsigned char* __restrict__ core, int* __restrict__ parent,
                                                       int* __restrict__ root) {
  const DgGrid gr = *grid;
  const float e2_hi = (float)eps2 * 1.0001f, e2_lo = (float)eps2 * 0.9999f;
  const int64_t p = (int64_t)blockIdx.x * DG_T + threadIdx.x;
  if (p >= n) return;
  const int i = idx_sorted[p];
  const bool core_i = phase > 0 && core[i];
  if (phase == 1 && !core_i) return;
  if (phase == 2 && core_i) {
    root[i] = dg_find(parent, i);
    return;
  }
  float xi[D];
#pragma unroll
  for (int c = 0; c < D; ++c) xi[c] = xs[p * D + c];
  int cell = keys_sorted[p];
  const int ic = cell % gr.g[2];
  cell /= gr.g[2];
  const int ib = cell % gr.g[1], ia = cell / gr.g[1];
  const int c_lo = max(ic - 1, 0), c_hi = min(ic + 1, gr.g[2] - 1);
  int count = 0, best = 0x7fffffff, my_root = i;
  for (int a = max(ia - 1, 0); a <= min(ia + 1, gr.g[0] - 1); ++a)
    for (int b = max(ib - 1, 0); b <= min(ib + 1, gr.g[1] - 1); ++b) {
      const int base = (a * gr.g[1] + b) * gr.g[2];
      const int q_end = cell_begin[base + c_hi + 1];
      for (int q = cell_begin[base + c_lo]; q < q_end; ++q) {
        if (phase == 0) {
          count += dg_near<D>(xi, xs + (int64_t)q * D, eps2, e2_hi, e2_lo) ? 1 : 0;
        } else {
          const int j = idx_sorted[q];
          if (phase == 1 ? (j >= i) : false) continue;  // pass 1: every pair once, from its higher index
          if (!core[j] || !dg_near<D>(xi, xs + (int64_t)q * D, eps2, e2_hi, e2_lo)) continue;
          if (phase == 1) {
            if (__ldcg(parent + j) != my_root) {
              dg_unite(parent, i, j);
              my_root = dg_find(parent, i);
            }
          } else {
            best = min(best, dg_find(parent, j));
          }
        }
      }
    }
  if (phase == 0) {
    core[i] = count >= min_pts ? 1 : 0;
    parent[i] = i;
  } else if (phase == 2) {
    root[i] = best == 0x7fffffff ? -1 : best;
  }
}

static size_t dg_align(size_t v) { return (v + 255) / 256 * 256; }

static size_t dg_cub_bytes(int64_t n) {
  size_t b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)n, 0, DG_CELL_BITS);
  return b;
}

size_t dbscan_grid_workspace_bytes(int64_t n) {
  if (n <= 0) return 256;
  return 256 + 256 + 4 * dg_align((size_t)n * 4) + dg_align(((size_t)1 << DG_CELL_BITS) * 4 + 8) +
         dg_align((size_t)n * DG_MAXD * 4) + dg_align(dg_cub_bytes(n));
}

struct DgWs {  // the workspace, carved
  unsigned* mm;
  DgGrid* grid;
  int32_t *keys, *vals, *keys_s, *idx_s, *cell_begin;
  float* xs;
  char* cub;
};

static DgWs dg_carve(char* ws, int64_t n) {
  DgWs w;
  w.mm = reinterpret_cast<unsigned*>(ws);
  w.grid = reinterpret_cast<DgGrid*>(ws + 256);
  char* p = ws + 512;
  w.keys = reinterpret_cast<int32_t*>(p); p += dg_align((size_t)n * 4);
  w.vals = reinterpret_cast<int32_t*>(p); p += dg_align((size_t)n * 4);
  w.keys_s = reinterpret_cast<int32_t*>(p); p += dg_align((size_t)n * 4);
  w.idx_s = reinterpret_cast<int32_t*>(p); p += dg_align((size_t)n * 4);
  w.cell_begin = reinterpret_cast<int32_t*>(p); p += dg_align(((size_t)1 << DG_CELL_BITS) * 4 + 8);
  w.xs = reinterpret_cast<float*>(p); p += dg_align((size_t)n * DG_MAXD * 4);
  w.cub = p;
  return w;
}

// the cell list of x for search radius eps, into the workspace
template <int D>
static int dg_build(const float* x, int d, int64_t n, double eps, const DgWs& w, cudaStream_t st) {
  size_t cub_bytes = dg_cub_bytes(n);
  const int k = d < 3 ? d : 3;
  const int threads = 256;
  const int blocks = (int)imin64((n + threads - 1) / threads, (int64_t)kNumSMs * 8);
  int rc = check_cuda(cudaMemsetAsync(w.mm, 0xff, 12, st), "cell list: init");  // running minima
  if (rc == GTB_OK) rc = check_cuda(cudaMemsetAsync(w.mm + 3, 0, 12, st), "cell list: init");  // running maxima
  if (rc) return rc;
  dg_bbox_kernel<<<blocks, threads, 0, st>>>(x, d, n, k, w.mm);
  GTB_CHECK_LAUNCH("dg_bbox_kernel");
  dg_setup_kernel<<<1, 32, 0, st>>>(w.mm, k, eps, w.grid);
  GTB_CHECK_LAUNCH("dg_setup_kernel");
  dg_keys_kernel<<<blocks, threads, 0, st>>>(x, d, n, w.grid, w.keys, w.vals);
  GTB_CHECK_LAUNCH("dg_keys_kernel");
  rc = check_cuda(cub::DeviceRadixSort::SortPairs(w.cub, cub_bytes, w.keys, w.keys_s, w.vals, w.idx_s, (int)n, 0, DG_CELL_BITS, st),
                  "cell list: SortPairs");
  if (rc) return rc;
  dg_layout_kernel<D><<<blocks, threads, 0, st>>>(x, d, n, w.grid, w.keys_s, w.idx_s, w.cell_begin, w.xs);
  GTB_CHECK_LAUNCH("dg_layout_kernel");
  return GTB_OK;
}

template <int D>
static int dg_run(const float* x, int d, int64_t n, double eps, int min_pts, unsigned char* core, int* parent, int* root,
                  char* ws, cudaStream_t st) {
  const DgWs w = dg_carve(ws, n);
  int rc = dg_build<D>(x, d, n, eps, w, st);
  if (rc) return rc;
  const double eps2 = eps * eps;
  const int pb = (int)((n + DG_T - 1) / DG_T);
  for (int phase = 0; phase < 3; ++phase) {
    dg_pass_kernel<D><<<pb, DG_T, 0, st>>>(w.xs, n, w.grid, w.keys_s, w.idx_s, w.cell_begin, eps2, min_pts, phase, core, parent, root);
    GTB_CHECK_LAUNCH("dg_pass_kernel");
  }
  return GTB_OK;
}

int dbscan_grid(const float* x, int d, int64_t n, double eps, int min_pts, unsigned char* core, int* parent, int* root,
                void* workspace, size_t workspace_bytes, cudaStream_t st) {
  GTB_REQUIRE(x && core && parent && root && d >= 1 && d <= DG_MAXD && n < (1ll << 31) - 1 && min_pts >= 1, GTB_ERR_BAD_ARG,
              "gtb_dbscan_grid_f32: bad arguments (dimension must be in [1, %d])", DG_MAXD);
  GTB_REQUIRE(workspace != nullptr && workspace_bytes >= dbscan_grid_workspace_bytes(n), GTB_ERR_WORKSPACE,
              "gtb_dbscan_grid_f32: workspace too small");
  GTB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, GTB_ERR_BAD_ARG, "gtb_dbscan_grid_f32: workspace must be 256-byte aligned");
  if (n == 0) return GTB_OK;
  char* ws = static_cast<char*>(workspace);
  if (d <= 4) return dg_run<4>(x, d, n, eps, min_pts, core, parent, root, ws, st);
  if (d <= 8) return dg_run<8>(x, d, n, eps, min_pts, core, parent, root, ws, st);
  return dg_run<16>(x, d, n, eps, min_pts, core, parent, root, ws, st);
}

// ---------------------------------------------------------------- radius graph over the cell list
// torch_cluster.radius_graph semantics as in radius.cu (strict fp32 ||x_i - x_j||^2 < r^2 accumulated over the
// coordinates in order, same batch entry, at most max_nb neighbours per centre: the lowest indices, ascending).
// The zero padding of the coordinates to D adds exact zeros to the sum, so the fp32 distance -- and with it
// the edge list -- is bit-identical to the all-pairs walk.  The cells are 1e-5 wider than r: a pair two cells
// apart differs by more than the cell width in one coordinate, which fp32 rounding (6e-8 relative per
// operation) cannot bring back below r^2.
// One thread per centre (in cell order).  FILL: the centre's segment of the edge list is kept sorted while
// the candidates arrive cell by cell (ascending inside a cell: the radix sort is stable), so most insertions
// are appends; over the cap a smaller index displaces the largest one.
template <int D, bool FILL>
__global__ void __launch_bounds__(DG_T) rg_pass_kernel(const float* __restrict__ xs, int64_t n, const DgGrid* __restrict__ grid,
                                                       const int32_t* __restrict__ keys_sorted,
                                                       const int32_t* __restrict__ idx_sorted,
                                                       const int32_t* __restrict__ cell_begin,
                                                       const int64_t* __restrict__ batch, float r2, int max_nb, int loop,
                                                       int32_t* __restrict__ counts, const int64_t* __restrict__ offsets,
                                                       int64_t* __restrict__ edge_index, int64_t n_edges) {
  const DgGrid gr = *grid;
  const int64_t p = (int64_t)blockIdx.x * DG_T + threadIdx.x;
  if (p >= n) return;
  const int i = idx_sorted[p];
  float xi[D];
#pragma unroll
  for (int c = 0; c < D; ++c) xi[c] = xs[p * D + c];
  const long long batch_i = batch ? batch[i] : 0;
  int64_t* seg = FILL ? edge_index + offsets[i] : nullptr;
  int cell = keys_sorted[p];
  const int ic = cell % gr.g[2];
  cell /= gr.g[2];
  const int ib = cell % gr.g[1], ia = cell / gr.g[1];
  const int c_lo = max(ic - 1, 0), c_hi = min(ic + 1, gr.g[2] - 1);
  int kept = 0;
  for (int a = max(ia - 1, 0); a <= min(ia + 1, gr.g[0] - 1); ++a)
    for (int b = max(ib - 1, 0); b <= min(ib + 1, gr.g[1] - 1); ++b) {
      const int base = (a * gr.g[1] + b) * gr.g[2];
      const int q_end = cell_begin[base + c_hi + 1];
      for (int q = cell_begin[base + c_lo]; q < q_end; ++q) {
        const float* xj = xs + (int64_t)q * D;
        float v[D];
#pragma unroll
        for (int w = 0; w < D / 4; ++w) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(xj) + w);
          v[4 * w] = t.x; v[4 * w + 1] = t.y; v[4 * w + 2] = t.z; v[4 * w + 3] = t.w;
        }
        float d2 = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const float t = xi[c] - v[c];
          d2 = fmaf(t, t, d2);
        }
        if (!(d2 < r2)) continue;
        const int j = idx_sorted[q];
        if (!loop && j == i) continue;
        if (batch && batch[j] != batch_i) continue;
        if (!FILL) {
          ++kept;
        } else if (kept < max_nb || (long long)j < seg[kept - 1]) {
          int pos = kept < max_nb ? kept++ : kept - 1;  // free slot at the end, or the largest kept index leaves
          while (pos > 0 && seg[pos - 1] > (long long)j) {
            seg[pos] = seg[pos - 1];
            --pos;
          }
          seg[pos] = j;
        }
      }
    }
  if (!FILL) {
    counts[i] = min(kept, max_nb);
  } else {
    int64_t* centre = seg + n_edges;
    for (int k = 0; k < kept; ++k) centre[k] = i;
  }
}

template <int D>
static int rg_count(const float* x, int d, int64_t n, const int64_t* batch, float r, int max_nb, int loop, int32_t* counts,
                    char* ws, cudaStream_t st) {
  const DgWs w = dg_carve(ws, n);
  int rc = dg_build<D>(x, d, n, fabs((double)r) * (1.0 + 1e-5), w, st);
  if (rc) return rc;
  rg_pass_kernel<D, false><<<(int)((n + DG_T - 1) / DG_T), DG_T, 0, st>>>(w.xs, n, w.grid, w.keys_s, w.idx_s, w.cell_begin, batch,
                                                                         r * r, max_nb, loop, counts, nullptr, nullptr, 0);
  GTB_CHECK_LAUNCH("rg_pass_kernel<count>");
  return GTB_OK;
}

template <int D>
static int rg_fill(int64_t n, const int64_t* batch, float r, int max_nb, int loop, const int64_t* offsets, int64_t* edge_index,
                   int64_t n_edges, char* ws, cudaStream_t st) {
  const DgWs w = dg_carve(ws, n);
  rg_pass_kernel<D, true><<<(int)((n + DG_T - 1) / DG_T), DG_T, 0, st>>>(w.xs, n, w.grid, w.keys_s, w.idx_s, w.cell_begin, batch,
                                                                        r * r, max_nb, loop, nullptr, offsets, edge_index, n_edges);
  GTB_CHECK_LAUNCH("rg_pass_kernel<fill>");
  return GTB_OK;
}

static int rg_check(const char* who, const void* x, int d, int64_t n, int max_nb, const void* workspace, size_t workspace_bytes) {
  GTB_REQUIRE(x && d >= 1 && d <= DG_MAXD && n < (1ll << 31) - 1 && max_nb >= 1, GTB_ERR_BAD_ARG,
              "%s: bad arguments (dimension must be in [1, %d])", who, DG_MAXD);
  GTB_REQUIRE(workspace != nullptr && workspace_bytes >= dbscan_grid_workspace_bytes(n), GTB_ERR_WORKSPACE, "%s: workspace too small", who);
  GTB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, GTB_ERR_BAD_ARG, "%s: workspace must be 256-byte aligned", who);
  return GTB_OK;
}

// builds the cell list in the workspace and counts; radius_graph_grid_fill walks the SAME workspace
int radius_graph_grid_count(const float* x, int d, int64_t n, const int64_t* batch, float r, int max_nb, int loop, int32_t* counts,
                            void* workspace, size_t workspace_bytes, cudaStream_t st) {
  int rc = rg_check("gtb_radius_graph_grid_count_f32", x, d, n, max_nb, workspace, workspace_bytes);
  if (rc) return rc;
  GTB_REQUIRE(counts != nullptr || n == 0, GTB_ERR_BAD_ARG, "gtb_radius_graph_grid_count_f32: counts is null");
  if (n == 0) return GTB_OK;
  char* ws = static_cast<char*>(workspace);
  if (d <= 4) return rg_count<4>(x, d, n, batch, r, max_nb, loop, counts, ws, st);
  if (d <= 8) return rg_count<8>(x, d, n, batch, r, max_nb, loop, counts, ws, st);
  return rg_count<16>(x, d, n, batch, r, max_nb, loop, counts, ws, st);
}

int radius_graph_grid_fill(const float* x, int d, int64_t n, const int64_t* batch, float r, int max_nb, int loop,
                           const int64_t* offsets, int64_t* edge_index, int64_t n_edges, void* workspace, size_t workspace_bytes,
                           cudaStream_t st) {
  int rc = rg_check("gtb_radius_graph_grid_fill_f32", x, d, n, max_nb, workspace, workspace_bytes);
  if (rc) return rc;
  GTB_REQUIRE(offsets && (edge_index || n_edges == 0), GTB_ERR_BAD_ARG, "gtb_radius_graph_grid_fill_f32: bad arguments");
  if (n == 0 || n_edges == 0) return GTB_OK;
  char* ws = static_cast<char*>(workspace);
  if (d <= 4) return rg_fill<4>(n, batch, r, max_nb, loop, offsets, edge_index, n_edges, ws, st);
  if (d <= 8) return rg_fill<8>(n, batch, r, max_nb, loop, offsets, edge_index, n_edges, ws, st);
  return rg_fill<16>(n, batch, r, max_nb, loop, offsets, edge_index, n_edges, ws, st);
}

// ---------------------------------------------------------------- fused pair sums over the cell list
// radius.cu's radius_pair_sum_kernel (the repulsive terms of the hinge and condensation losses, fused with the
// neighbour search: metric_learning.py:93-112, oc.py:46-69, 115-117) without the all-pairs walk.  Same edges
// (same fp32 distance test, same cap: the max_nb lowest neighbour indices of a centre), same fp32 terms; the
// float64 sums and the float atomics of the gradient add them in another order.
// The candidates of a centre arrive cell by cell, not in index order, so the cap is applied through a
// threshold: walk 1 counts the neighbours; only if they exceed the cap, a bisection over the index range
// (log2 n more walks, for that centre alone) finds the max_nb-th smallest neighbour index; walk 2 takes the
// neighbours up to it.
template <int D, class F>
__device__ __forceinline__ void rg_walk(const float* __restrict__ xs, const DgGrid& gr, const int32_t* __restrict__ cell_begin,
                                        int ia, int ib, int ic, const float (&xi)[D], float r2, F f) {
  const int c_lo = max(ic - 1, 0), c_hi = min(ic + 1, gr.g[2] - 1);
  for (int a = max(ia - 1, 0); a <= min(ia + 1, gr.g[0] - 1); ++a)
    for (int b = max(ib - 1, 0); b <= min(ib + 1, gr.g[1] - 1); ++b) {
      const int base = (a * gr.g[1] + b) * gr.g[2];
      const int q_end = cell_begin[base + c_hi + 1];
      for (int q = cell_begin[base + c_lo]; q < q_end; ++q) {
        const float* xj = xs + (int64_t)q * D;
        float v[D];
#pragma unroll
        for (int w = 0; w < D / 4; ++w) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(xj) + w);
          v[4 * w] = t.x; v[4 * w + 1] = t.y; v[4 * w + 2] = t.z; v[4 * w + 3] = t.w;
        }
        float d2 = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const float t = xi[c] - v[c];
          d2 = fmaf(t, t, d2);
        }
        if (d2 < r2) f(q, d2, v);
      }
    }
}

// modes, outputs and GRAD as radius_pair_sum_kernel (radius.cu)
template <int D, bool GRAD>
__global__ void __launch_bounds__(DG_T) rg_pair_sum_kernel(
    const float* __restrict__ xs, int d, int64_t n, const DgGrid* __restrict__ grid, const int32_t* __restrict__ keys_sorted,
    const int32_t* __restrict__ idx_sorted, const int32_t* __restrict__ cell_begin, const int64_t* __restrict__ batch,
    const int64_t* __restrict__ pid, const unsigned char* __restrict__ src_flag, const float* __restrict__ beta, float q_min,
    float r, float p, float eps, int max_nb, int mode, double* __restrict__ out, const float* __restrict__ coef,
    float* __restrict__ gx, float* __restrict__ gq) {
  __shared__ double red[4][DG_T / 32];
  const DgGrid gr = *grid;
  const int tid = threadIdx.x;
  const float r2 = r * r;
  const float c0 = GRAD ? __ldg(coef) : 0.f;
  double acc = 0.0, cnt_e = 0.0, nsum = 0.0, ncnt = 0.0;
  const int64_t pos = (int64_t)blockIdx.x * DG_T + tid;
  if (pos < n) {
    const int i = idx_sorted[pos];
    float xi[D], gi[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
      xi[c] = xs[pos * D + c];
      gi[c] = 0.f;
    }
    float gqi = 0.f;
    const long long pid_i = pid[i], batch_i = batch ? batch[i] : 0;
    float q_i = 0.f;
    if (beta) {
      const float b = __ldg(beta + i), a = atanhf(b);
      q_i = a * a + q_min;
      if (pid_i == 0) {
        nsum += (double)b;
        ncnt += 1.0;
      }
    }
    int cell = keys_sorted[pos];
    const int ic = cell % gr.g[2];
    cell /= gr.g[2];
    const int ib = cell % gr.g[1], ia = cell / gr.g[1];
    int n_nb = 0;
    rg_walk<D>(xs, gr, cell_begin, ia, ib, ic, xi, r2, [&](int q, float, const float (&)[D]) {
      const int j = idx_sorted[q];
      if (j != i && (!batch || batch[j] == batch_i)) ++n_nb;
    });
    int thr = 0x7fffffff;
    if (n_nb > max_nb) {  // smallest T with #{neighbours j <= T} >= max_nb
      int lo = 0, hi = (int)n - 1;
      while (lo < hi) {
        const int mid = lo + (hi - lo) / 2;
        int c = 0;
        rg_walk<D>(xs, gr, cell_begin, ia, ib, ic, xi, r2, [&](int q, float, const float (&)[D]) {
          const int j = idx_sorted[q];
          if (j <= mid && j != i && (!batch || batch[j] == batch_i)) ++c;
        });
        if (c >= max_nb) hi = mid;
        else lo = mid + 1;
      }
      thr = lo;
    }
    if (n_nb > 0)
      rg_walk<D>(xs, gr, cell_begin, ia, ib, ic, xi, r2, [&](int q, float d2, const float (&v)[D]) {
        const int j = idx_sorted[q];
        if (j > thr || j == i || (batch && batch[j] != batch_i)) return;
        if (!src_flag[j] || pid[j] == pid_i) return;
        float q_j = 0.f;
        if (beta) {
          const float a = atanhf(__ldg(beta + j));
          q_j = a * a + q_min;
        }
        if (!GRAD) {
          float term;
          if (mode == 0) term = fmaxf(r - powf(sqrtf(d2), p), 0.f);
          else term = (r - sqrtf(eps + d2)) * q_j * q_i;
          acc += (double)term;
          cnt_e += 1.0;
        } else {
          float w = 0.f;
          if (mode == 0) {
            const float dist = sqrtf(d2);
            if (dist > 0.f && r - powf(dist, p) > 0.f) w = -c0 * p * powf(dist, p - 2.f);
          } else {
            const float s = sqrtf(eps + d2), qq = q_j * q_i;
            w = -c0 * qq / s;
            gqi += c0 * (r - s) * q_j;
            atomicAdd(gq + j, c0 * (r - s) * q_i);
          }
          if (w != 0.f) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
              if (c < d) {
                const float g = w * (xi[c] - v[c]);
                gi[c] += g;
                atomicAdd(gx + (size_t)j * d + c, -g);
              }
            }
          }
        }
      });
    if (GRAD) {
#pragma unroll
      for (int c = 0; c < D; ++c)
        if (c < d && gi[c] != 0.f) atomicAdd(gx + (size_t)i * d + c, gi[c]);
      if (mode == 1 && gqi != 0.f) atomicAdd(gq + i, gqi);
    }
  }
  if (GRAD) return;
  double v[4] = {acc, cnt_e, nsum, ncnt};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((tid & 31) == 0) red[k][tid >> 5] = v[k];
  }
  __syncthreads();
  if (tid < 4) {
    double s = 0.0;
    for (int w = 0; w < DG_T / 32; ++w) s += red[tid][w];
    if (s != 0.0) atomicAdd(out + tid, s);
  }
}

template <int D>
static int rg_pair(const float* x, int d, int64_t n, const int64_t* batch, const int64_t* pid, const unsigned char* src_flag,
                   const float* beta, float q_min, float r, float p, float eps, int max_nb, int mode, double* out,
                   const float* coef, float* gx, float* gq, char* ws, cudaStream_t st) {
  const DgWs w = dg_carve(ws, n);
  int rc = dg_build<D>(x, d, n, fabs((double)r) * (1.0 + 1e-5), w, st);
  if (rc) return rc;
  const int blocks = (int)((n + DG_T - 1) / DG_T);
  if (coef == nullptr)
    rg_pair_sum_kernel<D, false><<<blocks, DG_T, 0, st>>>(w.xs, d, n, w.grid, w.keys_s, w.idx_s, w.cell_begin, batch, pid, src_flag,
                                                          beta, q_min, r, p, eps, max_nb, mode, out, nullptr, nullptr, nullptr);
  else
    rg_pair_sum_kernel<D, true><<<blocks, DG_T, 0, st>>>(w.xs, d, n, w.grid, w.keys_s, w.idx_s, w.cell_begin, batch, pid, src_flag,
                                                         beta, q_min, r, p, eps, max_nb, mode, nullptr, coef, gx, gq);
  GTB_CHECK_LAUNCH("rg_pair_sum_kernel");
  return GTB_OK;
}

// forward (coef == nullptr: out += the four sums) or gradient (coef != nullptr: gx, gq += the gradient), see
// radius_pair_sum / radius_pair_sum_grad in radius.cu
int radius_pair_sum_grid(const float* x, int d, int64_t n, const int64_t* batch, const int64_t* pid, const unsigned char* src_flag,
                         const float* beta, float q_min, float r, float p, float eps, int max_nb, int mode, double* out,
                         const float* coef, float* gx, float* gq, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const char* who = coef ? "gtb_radius_pair_sum_grad_grid_f32" : "gtb_radius_pair_sum_grid_f32";
  int rc = rg_check(who, x, d, n, max_nb, workspace, workspace_bytes);
  if (rc) return rc;
  GTB_REQUIRE(pid && src_flag && (mode == 0 || (mode == 1 && beta != nullptr)) && (coef ? (gx && (mode == 0 || gq)) : out != nullptr),
              GTB_ERR_BAD_ARG, "%s: bad arguments", who);
  if (n == 0) return GTB_OK;
  char* ws = static_cast<char*>(workspace);
  if (d <= 4) return rg_pair<4>(x, d, n, batch, pid, src_flag, beta, q_min, r, p, eps, max_nb, mode, out, coef, gx, gq, ws, st);
  if (d <= 8) return rg_pair<8>(x, d, n, batch, pid, src_flag, beta, q_min, r, p, eps, max_nb, mode, out, coef, gx, gq, ws, st);
  return rg_pair<16>(x, d, n, batch, pid, src_flag, beta, q_min, r, p, eps, max_nb, mode, out, coef, gx, gq, ws, st);
}

}  // namespace gtb
