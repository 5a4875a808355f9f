// Fixed-radius pair sums in a low-dimensional latent space: the native work behind the
// reference's torch_cluster.radius_graph call sites (metrics/losses/oc.py:115-117,
// metrics/losses/metric_learning.py:97-103), fused with the potentials that consume the
// edges, so the edge list is never materialised.
//
// torch_cluster.radius_graph(x, r, batch, loop=False, max_num_neighbors) semantics kept
// (restated in oracle/losses_oracle.py:radius_graph): an edge (neighbour j -> centre i) exists
// when ||x_i - x_j||^2 < r^2 (strict), i != j, same batch entry; a centre keeps at most
// max_num_neighbors neighbours -- here, as in the oracle, the ones with the LOWEST indices
// (torch_cluster's choice is implementation-defined; untruncated graphs are identical).
//
// One thread per centre walks all candidate neighbours in ascending order (tiles of 256 hits
// staged in shared memory).  N^2 * d multiply-adds: 1e10 pair checks at N = 1e5, a few ms.
#include "common.cuh"

namespace gtb {

constexpr int RAD_T = 256;      // threads = hits per staged tile
constexpr int RAD_MAXD = 16;    // latent dimensions

struct RadTile {
  float x[RAD_MAXD][RAD_T];     // transposed: conflict-free broadcast reads of one neighbour
  long long pid[RAD_T];
  long long batch[RAD_T];
  float q[RAD_T];
  unsigned char flag[RAD_T];
};

// mode 0 (hinge repulsion, metric_learning.py:93-112, 47-52): edge kept when src_flag[j] and
//        pid[j] != pid[i]; term = relu(r - dist^p)
// mode 1 (condensation repulsion, oc.py:46-69): edge kept when src_flag[j] (j is a condensation
//        point) and pid[j] != pid[i]; term = (r - sqrt(eps + dist^2)) * q_j * q_i,
//        q = atanh(beta)^2 + q_min
// out: {sum of terms, number of kept edges, sum of beta over pid == 0, number of pid == 0}
// GRAD: instead of the sums, the gradient of `coef[0] * (sum of terms)` w.r.t. x (gx, zero-filled by
// the caller) and, in mode 1, w.r.t. the charges (gq, zero-filled): the centre's share is kept in
// registers, the neighbour's share goes out through atomics.  d dist / d x = 0 at dist = 0, as
// torch.norm's backward.
template <bool GRAD>
__global__ void __launch_bounds__(RAD_T) radius_pair_sum_kernel(
    const float* __restrict__ x, int d, int64_t n, const int64_t* __restrict__ batch,
    const int64_t* __restrict__ pid, const unsigned char* __restrict__ src_flag, const float* __restrict__ beta,
    float q_min, float r, float p, float eps, int max_nb, int mode, double* __restrict__ out,
    const float* __restrict__ coef, float* __restrict__ gx, float* __restrict__ gq) {
  __shared__ RadTile tile;
  __shared__ double red[4][RAD_T / 32];
  const int tid = threadIdx.x;
  const float r2 = r * r;
  const float c0 = GRAD ? __ldg(coef) : 0.f;
  double acc = 0.0, cnt_e = 0.0, nsum = 0.0, ncnt = 0.0;
  const int64_t n_round = (n + RAD_T - 1) / RAD_T * RAD_T;
  for (int64_t i0 = (int64_t)blockIdx.x * RAD_T; i0 < n_round; i0 += (int64_t)gridDim.x * RAD_T) {
    const int64_t i = i0 + tid;
    const bool have = i < n;
    float xi[RAD_MAXD], gi[RAD_MAXD];
#pragma unroll
    for (int c = 0; c < RAD_MAXD; ++c) {
      xi[c] = (have && c < d) ? __ldg(x + (size_t)i * d + c) : 0.f;
      gi[c] = 0.f;
    }
    float gqi = 0.f;
    const long long pid_i = have ? pid[i] : 0, batch_i = (have && batch) ? batch[i] : 0;
    float q_i = 0.f;
    if (have && beta) {
      const float b = __ldg(beta + i), a = atanhf(b);
      q_i = a * a + q_min;
      if (pid_i == 0) {
        nsum += (double)b;
        ncnt += 1.0;
      }
    }
    int kept = 0;
    for (int64_t j0 = 0; j0 < n; j0 += RAD_T) {
      __syncthreads();
      {
        const int64_t j = j0 + tid;
        const bool hj = j < n;
        for (int c = 0; c < d; ++c) tile.x[c][tid] = hj ? __ldg(x + (size_t)j * d + c) : 0.f;
        tile.pid[tid] = hj ? pid[j] : 0;
        tile.batch[tid] = (hj && batch) ? batch[j] : 0;
        tile.flag[tid] = hj ? src_flag[j] : 0;
        float qj = 0.f;
        if (hj && beta) {
          const float a = atanhf(__ldg(beta + j));
          qj = a * a + q_min;
        }
        tile.q[tid] = qj;
      }
      __syncthreads();
      if (!have || kept >= max_nb) continue;
      const int lim = (int)min((int64_t)RAD_T, n - j0);
      for (int jj = 0; jj < lim; ++jj) {
        float d2 = 0.f;
#pragma unroll
        for (int c = 0; c < RAD_MAXD; ++c) {
          if (c < d) {
            const float t = xi[c] - tile.x[c][jj];
            d2 = fmaf(t, t, d2);
          }
        }
        if (d2 < r2 && j0 + jj != i && tile.batch[jj] == batch_i) {
          if (kept >= max_nb) break;
          ++kept;
          if (tile.flag[jj] && tile.pid[jj] != pid_i) {
            if (!GRAD) {
              float term;
              if (mode == 0) term = fmaxf(r - powf(sqrtf(d2), p), 0.f);
              else term = (r - sqrtf(eps + d2)) * tile.q[jj] * q_i;
              acc += (double)term;
              cnt_e += 1.0;
            } else {
              // term = f(dist): d term / d x_i = w (x_i - x_j), the neighbour gets the opposite
              float w = 0.f;
              if (mode == 0) {
                const float dist = sqrtf(d2);
                if (dist > 0.f && r - powf(dist, p) > 0.f) w = -c0 * p * powf(dist, p - 2.f);
              } else {
                const float s = sqrtf(eps + d2), qq = tile.q[jj] * q_i;
                w = -c0 * qq / s;
                gqi += c0 * (r - s) * tile.q[jj];
                atomicAdd(gq + j0 + jj, c0 * (r - s) * q_i);
              }
              if (w != 0.f) {
#pragma unroll
                for (int c = 0; c < RAD_MAXD; ++c) {
                  if (c < d) {
                    const float g = w * (xi[c] - tile.x[c][jj]);
                    gi[c] += g;
                    atomicAdd(gx + (size_t)(j0 + jj) * d + c, -g);
                  }
                }
              }
            }
          }
        }
      }
    }
    if (GRAD && have) {
      for (int c = 0; c < d; ++c)
        if (gi[c] != 0.f) atomicAdd(gx + (size_t)i * d + c, gi[c]);
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
    for (int w = 0; w < RAD_T / 32; ++w) s += red[tid][w];
    atomicAdd(out + tid, s);
  }
}

int radius_pair_sum(const float* x, int d, int64_t n, const int64_t* batch, const int64_t* pid,
                    const unsigned char* src_flag, const float* beta, float q_min, float r, float p, float eps,
                    int max_nb, int mode, double* out, cudaStream_t st) {
  GTB_REQUIRE(x && pid && src_flag && out && d >= 1 && d <= RAD_MAXD && (mode == 0 || (mode == 1 && beta != nullptr)) &&
                  max_nb >= 1,
              GTB_ERR_BAD_ARG, "gtb_radius_pair_sum_f32: bad arguments (latent dimension must be in [1, %d])", RAD_MAXD);
  if (n == 0) return GTB_OK;
  const int blocks = (int)imin64((n + RAD_T - 1) / RAD_T, (int64_t)kNumSMs * 4);
  radius_pair_sum_kernel<false><<<blocks, RAD_T, 0, st>>>(x, d, n, batch, pid, src_flag, beta, q_min, r, p, eps, max_nb, mode,
                                                          out, nullptr, nullptr, nullptr);
  GTB_CHECK_LAUNCH("radius_pair_sum_kernel");
  return GTB_OK;
}

int radius_pair_sum_grad(const float* x, int d, int64_t n, const int64_t* batch, const int64_t* pid,
                         const unsigned char* src_flag, const float* beta, float q_min, float r, float p, float eps,
                         int max_nb, int mode, const float* coef, float* gx, float* gq, cudaStream_t st) {
  GTB_REQUIRE(x && pid && src_flag && coef && gx && d >= 1 && d <= RAD_MAXD &&
                  (mode == 0 || (mode == 1 && beta != nullptr && gq != nullptr)) && max_nb >= 1,
              GTB_ERR_BAD_ARG, "gtb_radius_pair_sum_grad_f32: bad arguments (latent dimension must be in [1, %d])", RAD_MAXD);
  if (n == 0) return GTB_OK;
  const int blocks = (int)imin64((n + RAD_T - 1) / RAD_T, (int64_t)kNumSMs * 4);
  radius_pair_sum_kernel<true><<<blocks, RAD_T, 0, st>>>(x, d, n, batch, pid, src_flag, beta, q_min, r, p, eps, max_nb, mode,
                                                         nullptr, coef, gx, gq);
  GTB_CHECK_LAUNCH("radius_pair_sum_kernel<grad>");
  return GTB_OK;
}

// sum over the edges whose SOURCE is flagged of ||x_a - x_b||^p, and their number: the attractive
// term of the hinge loss (metric_learning.py:28-30, 111)
__global__ void edge_dist_pow_sum_kernel(const float* __restrict__ x, int d, const int64_t* __restrict__ edges,
                                         int64_t n_edges, const unsigned char* __restrict__ src_flag, float p,
                                         double* __restrict__ out) {
  double acc = 0.0, cnt = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += stride) {
    const int64_t a = edges[e], b = edges[n_edges + e];
    if (!src_flag[a]) continue;
    float d2 = 0.f;
    for (int c = 0; c < d; ++c) {
      const float t = x[(size_t)a * d + c] - x[(size_t)b * d + c];
      d2 = fmaf(t, t, d2);
    }
    acc += (double)powf(sqrtf(d2), p);
    cnt += 1.0;
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, acc);
    atomicAdd(out + 1, cnt);
  }
}

// gradient of coef[0] * (the sum above) w.r.t. x, added onto gx
__global__ void edge_dist_pow_grad_kernel(const float* __restrict__ x, int d, const int64_t* __restrict__ edges,
                                          int64_t n_edges, const unsigned char* __restrict__ src_flag, float p,
                                          const float* __restrict__ coef, float* __restrict__ gx) {
  const float c0 = __ldg(coef);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += stride) {
    const int64_t a = edges[e], b = edges[n_edges + e];
    if (!src_flag[a]) continue;
    float d2 = 0.f;
    for (int c = 0; c < d; ++c) {
      const float t = x[(size_t)a * d + c] - x[(size_t)b * d + c];
      d2 = fmaf(t, t, d2);
    }
    const float dist = sqrtf(d2);
    if (!(dist > 0.f)) continue;
    const float w = c0 * p * powf(dist, p - 2.f);
    for (int c = 0; c < d; ++c) {
      const float g = w * (x[(size_t)a * d + c] - x[(size_t)b * d + c]);
      atomicAdd(gx + (size_t)a * d + c, g);
      atomicAdd(gx + (size_t)b * d + c, -g);
    }
  }
}

int edge_dist_pow_grad(const float* x, int d, const int64_t* edges, int64_t n_edges, const unsigned char* src_flag,
                       float p, const float* coef, float* gx, cudaStream_t st) {
  GTB_REQUIRE(x && src_flag && coef && gx && d >= 1 && (edges || n_edges == 0), GTB_ERR_BAD_ARG,
              "gtb_edge_dist_pow_grad_f32: bad arguments");
  if (n_edges == 0) return GTB_OK;
  const int blocks = (int)imin64((n_edges + 255) / 256, (int64_t)kNumSMs * 8);
  edge_dist_pow_grad_kernel<<<blocks, 256, 0, st>>>(x, d, edges, n_edges, src_flag, p, coef, gx);
  GTB_CHECK_LAUNCH("edge_dist_pow_grad_kernel");
  return GTB_OK;
}

int edge_dist_pow_sum(const float* x, int d, const int64_t* edges, int64_t n_edges, const unsigned char* src_flag,
                      float p, double* out, cudaStream_t st) {
  GTB_REQUIRE(x && src_flag && out && d >= 1 && (edges || n_edges == 0), GTB_ERR_BAD_ARG, "gtb_edge_dist_pow_sum_f32: bad arguments");
  if (n_edges == 0) return GTB_OK;
  const int blocks = (int)imin64((n_edges + 255) / 256, (int64_t)kNumSMs * 8);
  edge_dist_pow_sum_kernel<<<blocks, 256, 0, st>>>(x, d, edges, n_edges, src_flag, p, out);
  GTB_CHECK_LAUNCH("edge_dist_pow_sum_kernel");
  return GTB_OK;
}

// ---------------------------------------------------------------- materialised radius graph
// torch_cluster.radius_graph as an edge list for callers outside the fused losses: pass 1 counts the
// kept neighbours of every centre, the host prefix-sums the counts, pass 2 writes the edges grouped
// by centre, neighbours ascending: row 0 = neighbour (source), row 1 = centre (target).
template <bool FILL>
__global__ void __launch_bounds__(RAD_T) radius_graph_kernel(const float* __restrict__ x, int d, int64_t n,
                                                             const int64_t* __restrict__ batch, float r, int max_nb,
                                                             int loop, int32_t* __restrict__ counts,
                                                             const int64_t* __restrict__ offsets,
                                                             int64_t* __restrict__ edge_index, int64_t n_edges) {
  __shared__ float tx[RAD_MAXD][RAD_T];
  __shared__ long long tb[RAD_T];
  const int tid = threadIdx.x;
  const float r2 = r * r;
  const int64_t n_round = (n + RAD_T - 1) / RAD_T * RAD_T;
  for (int64_t i0 = (int64_t)blockIdx.x * RAD_T; i0 < n_round; i0 += (int64_t)gridDim.x * RAD_T) {
    const int64_t i = i0 + tid;
    const bool have = i < n;
    float xi[RAD_MAXD];
#pragma unroll
    for (int c = 0; c < RAD_MAXD; ++c) xi[c] = (have && c < d) ? __ldg(x + (size_t)i * d + c) : 0.f;
    const long long batch_i = (have && batch) ? batch[i] : 0;
    const int64_t off = (FILL && have) ? offsets[i] : 0;
    int kept = 0;
    for (int64_t j0 = 0; j0 < n; j0 += RAD_T) {
      __syncthreads();
      {
        const int64_t j = j0 + tid;
        const bool hj = j < n;
        for (int c = 0; c < d; ++c) tx[c][tid] = hj ? __ldg(x + (size_t)j * d + c) : 0.f;
        tb[tid] = (hj && batch) ? batch[j] : 0;
      }
      __syncthreads();
      if (!have || kept >= max_nb) continue;
      const int lim = (int)min((int64_t)RAD_T, n - j0);
      for (int jj = 0; jj < lim && kept < max_nb; ++jj) {
        float d2 = 0.f;
#pragma unroll
        for (int c = 0; c < RAD_MAXD; ++c) {
          if (c < d) {
            const float t = xi[c] - tx[c][jj];
            d2 = fmaf(t, t, d2);
          }
        }
        if (d2 < r2 && (loop || j0 + jj != i) && tb[jj] == batch_i) {
          if (FILL) {
            edge_index[off + kept] = j0 + jj;
            edge_index[n_edges + off + kept] = i;
          }
          ++kept;
        }
      }
    }
    if (!FILL && have) counts[i] = kept;
  }
}

int radius_graph_count(const float* x, int d, int64_t n, const int64_t* batch, float r, int max_nb, int loop,
                       int32_t* counts, cudaStream_t st) {
  GTB_REQUIRE(x && counts && d >= 1 && d <= RAD_MAXD && max_nb >= 1, GTB_ERR_BAD_ARG,
              "gtb_radius_graph_count_f32: bad arguments (dimension must be in [1, %d])", RAD_MAXD);
  if (n == 0) return GTB_OK;
  const int blocks = (int)imin64((n + RAD_T - 1) / RAD_T, (int64_t)kNumSMs * 4);
  radius_graph_kernel<false><<<blocks, RAD_T, 0, st>>>(x, d, n, batch, r, max_nb, loop, counts, nullptr, nullptr, 0);
  GTB_CHECK_LAUNCH("radius_graph_kernel<count>");
  return GTB_OK;
}

int radius_graph_fill(const float* x, int d, int64_t n, const int64_t* batch, float r, int max_nb, int loop,
                      const int64_t* offsets, int64_t* edge_index, int64_t n_edges, cudaStream_t st) {
  GTB_REQUIRE(x && offsets && (edge_index || n_edges == 0) && d >= 1 && d <= RAD_MAXD && max_nb >= 1, GTB_ERR_BAD_ARG,
              "gtb_radius_graph_fill_f32: bad arguments (dimension must be in [1, %d])", RAD_MAXD);
  if (n == 0 || n_edges == 0) return GTB_OK;
  const int blocks = (int)imin64((n + RAD_T - 1) / RAD_T, (int64_t)kNumSMs * 4);
  radius_graph_kernel<true><<<blocks, RAD_T, 0, st>>>(x, d, n, batch, r, max_nb, loop, nullptr, offsets, edge_index, n_edges);
  GTB_CHECK_LAUNCH("radius_graph_kernel<fill>");
  return GTB_OK;
}

}  // namespace gtb
