// DBSCAN on the GPU for the hyper-parameter scan of the clustered latent space (reference
// postprocessing/fastrescanner.py:6-66 -> sklearn radius_neighbors + dbscan_inner;
// postprocessing/dbscanscanner.py:146-187).  Labels are IDENTICAL to sklearn's, not just up to
// relabelling:
//   * neighbourhood: dist <= eps (non-strict, the point itself included), evaluated in float64 like
//     sklearn does for float32 inputs;
//   * core sample: at least min_pts neighbours;
//   * dbscan_inner numbers clusters in the order of their lowest-index core sample and gives a
//     border point to the first cluster that reaches it, i.e. the lowest-numbered adjacent one:
//     here every component of core samples is rooted at its lowest index (lock-free union-find,
//     larger root hooked under the smaller), a border point takes the smallest adjacent root, and
//     the host numbers the roots in increasing order.
// Brute force over shared-memory tiles of candidates (latent spaces of the path have 2-8 dims).
#include "common.cuh"

namespace gtb {

constexpr int DB_T = 256;
constexpr int DB_MAXD = 16;

// candidate tile: coordinates zero-padded to D (a multiple of 4) so that a pair costs D/4 broadcast
// LDS.128 and no predicates
template <int D>
struct DbTile {
  float4 x[DB_T][D / 4];
  unsigned char core[DB_T];
};

template <int D>
__device__ __forceinline__ bool db_near(const float (&xi)[D], const DbTile<D>& t, int jj, double eps2, float e2_hi,
                                        float e2_lo) {
  // fp32 screen (relative error of the sum < D * 2^-23, far inside the 1e-4 margins); only pairs
  // on the rim of the eps-ball pay for the float64 evaluation that decides like sklearn
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < D / 4; ++q) {
    const float4 v = t.x[jj][q];
    const float u0 = xi[4 * q] - v.x, u1 = xi[4 * q + 1] - v.y, u2 = xi[4 * q + 2] - v.z, u3 = xi[4 * q + 3] - v.w;
    s = fmaf(u0, u0, s);
    s = fmaf(u1, u1, s);
    s = fmaf(u2, u2, s);
    s = fmaf(u3, u3, s);
  }
  if (s > e2_hi) return false;
  if (s < e2_lo) return true;
  double d2 = 0.0;
#pragma unroll
  for (int q = 0; q < D / 4; ++q) {
    const float4 v = t.x[jj][q];
    const double u0 = (double)xi[4 * q] - (double)v.x, u1 = (double)xi[4 * q + 1] - (double)v.y;
    const double u2 = (double)xi[4 * q + 2] - (double)v.z, u3 = (double)xi[4 * q + 3] - (double)v.w;
    d2 = fma(u0, u0, d2);
    d2 = fma(u1, u1, d2);
    d2 = fma(u2, u2, d2);
    d2 = fma(u3, u3, d2);
  }
  return d2 <= eps2;
}

__device__ __forceinline__ int db_find(int* parent, int v) {
  while (true) {
    // L2-coherent accesses: a stale L1 line would make a failed CAS retry forever
    const int p = __ldcg(parent + v);
    if (p == v) return v;
    const int gp = __ldcg(parent + p);
    if (gp != p) __stcg(parent + v, gp);  // path halving; racing writers only ever store ancestors
    v = p;
  }
}

__device__ __forceinline__ void db_unite(int* parent, int a, int b) {
  while (true) {
    a = db_find(parent, a);
    b = db_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    if (atomicCAS(parent + a, a, b) == a) return;  // hook the larger root under the smaller one
  }
}

// phase 0: neighbour counts -> core flags, parent[i] = i
// phase 1: unite every core sample with its lower-index core neighbours
// phase 2: root of every core sample; smallest adjacent root for the others (-1: noise)
template <int D>
__global__ void __launch_bounds__(DB_T) dbscan_kernel(const float* __restrict__ x, int d, int64_t n, double eps2,
                                                      int min_pts, int phase, unsigned char* __restrict__ core,
                                                      int* __restrict__ parent, int* __restrict__ root) {
  __shared__ DbTile<D> tile;
  const int tid = threadIdx.x;
  const int64_t n_tiles = (n + DB_T - 1) / DB_T;
  const float e2_hi = (float)eps2 * 1.0001f, e2_lo = (float)eps2 * 0.9999f;
  // the triangular phase 1 hands out its longest rows first
  for (int64_t it = blockIdx.x; it < n_tiles; it += gridDim.x) {
    const int64_t i0 = (phase == 1 ? n_tiles - 1 - it : it) * DB_T;
    const int64_t i = i0 + tid;
    const bool have = i < n;
    float xi[D];
#pragma unroll
    for (int c = 0; c < D; ++c) xi[c] = (have && c < d) ? __ldg(x + (size_t)i * d + c) : 0.f;
    const bool core_i = have && phase > 0 && core[i];
    int count = 0, best = 0x7fffffff;
    int my_root = (phase == 1 && core_i) ? (int)i : -1;
    const int64_t j_end = min((phase == 1) ? i0 + DB_T : n, n);
    for (int64_t j0 = 0; j0 < j_end; j0 += DB_T) {
      __syncthreads();
      {
        const int64_t j = j0 + tid;
        const bool hj = j < n;
        float v[D];
#pragma unroll
        for (int c = 0; c < D; ++c) v[c] = (hj && c < d) ? __ldg(x + (size_t)j * d + c) : 0.f;
#pragma unroll
        for (int q = 0; q < D / 4; ++q) tile.x[tid][q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        tile.core[tid] = (hj && phase > 0) ? core[j] : 0;
      }
      __syncthreads();
      if (!have) continue;
      const int lim = (int)min((int64_t)DB_T, n - j0);
      if (phase == 0) {
#pragma unroll 4
        for (int jj = 0; jj < lim; ++jj) count += db_near<D>(xi, tile, jj, eps2, e2_hi, e2_lo) ? 1 : 0;
      } else if (phase == 1) {
        if (!core_i) continue;
        const int lim1 = (int)min((int64_t)lim, i - j0);
        for (int jj = 0; jj < lim1; ++jj)
          if (tile.core[jj] && db_near<D>(xi, tile, jj, eps2, e2_hi, e2_lo)) {
            // one L2 read settles the common case (neighbour already under this thread's root)
            const int j = (int)(j0 + jj);
            if (__ldcg(parent + j) != my_root) {
              db_unite(parent, (int)i, j);
              my_root = db_find(parent, (int)i);
            }
          }
      } else {
        if (core_i) continue;
        for (int jj = 0; jj < lim; ++jj)
          if (tile.core[jj] && db_near<D>(xi, tile, jj, eps2, e2_hi, e2_lo)) best = min(best, db_find(parent, (int)(j0 + jj)));
      }
    }
    if (!have) continue;
    if (phase == 0) {
      core[i] = count >= min_pts ? 1 : 0;
      parent[i] = (int)i;
    } else if (phase == 2) {
      root[i] = core_i ? db_find(parent, (int)i) : (best == 0x7fffffff ? -1 : best);
    }
  }
}

template <int D>
static void db_launch(int blocks, cudaStream_t st, const float* x, int d, int64_t n, double eps2, int min_pts, int phase,
                      unsigned char* core, int* parent, int* root) {
  dbscan_kernel<D><<<blocks, DB_T, 0, st>>>(x, d, n, eps2, min_pts, phase, core, parent, root);
}

int dbscan(const float* x, int d, int64_t n, double eps, int min_pts, unsigned char* core, int* parent, int* root,
           cudaStream_t st) {
  GTB_REQUIRE(x && core && parent && root && d >= 1 && d <= DB_MAXD && n < (1ll << 31) - 1 && min_pts >= 1, GTB_ERR_BAD_ARG,
              "gtb_dbscan_f32: bad arguments (dimension must be in [1, %d])", DB_MAXD);
  if (n == 0) return GTB_OK;
  const double eps2 = eps * eps;
  const int blocks = (int)imin64((n + DB_T - 1) / DB_T, (int64_t)kNumSMs * 4);
  for (int phase = 0; phase < 3; ++phase) {
    if (d <= 4) db_launch<4>(blocks, st, x, d, n, eps2, min_pts, phase, core, parent, root);
    else if (d <= 8) db_launch<8>(blocks, st, x, d, n, eps2, min_pts, phase, core, parent, root);
    else db_launch<16>(blocks, st, x, d, n, eps2, min_pts, phase, core, parent, root);
    GTB_CHECK_LAUNCH("dbscan_kernel");
  }
  return GTB_OK;
}

}  // namespace gtb
