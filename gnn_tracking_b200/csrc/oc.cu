// Object-condensation loss "tiger" (reference metrics/losses/oc.py:251-347) without the
// N x K planes: ids are uniqued on the device, condensation points found by a packed
// atomic arg-max, and the attractive / repulsive potentials accumulated by a tiled
// hits x condensation-points kernel that keeps the CP tile in shared memory.
#include <cub/cub.cuh>

#include "common.cuh"

namespace gtb {

constexpr int64_t kSentinel = INT64_MAX;
constexpr int OC_MAXD = 16;

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

__global__ void oc_keys_kernel(const int64_t* __restrict__ id, const uint8_t* __restrict__ mask, int64_t n,
                               int64_t* __restrict__ keys) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
    keys[j] = mask[j] ? id[j] : kSentinel;
}

__global__ void oc_slots_kernel(const int64_t* __restrict__ id, int64_t n, const int64_t* __restrict__ uniq,
                                const int32_t* __restrict__ n_sel, int32_t* __restrict__ n_uniq,
                                int32_t* __restrict__ slot) {
  int k = *n_sel;
  if (k > 0 && uniq[k - 1] == kSentinel) --k;
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_uniq = k;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const int64_t v = id[j];
    int lo = 0, hi = k;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (uniq[mid] < v) lo = mid + 1; else hi = mid;
    }
    slot[j] = (lo < k && uniq[lo] == v) ? lo : -1;
  }
}

size_t oc_workspace_bytes(int64_t n) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, a, (const int64_t*)nullptr, (int64_t*)nullptr, (int)n);
  cub::DeviceSelect::Unique(nullptr, b, (const int64_t*)nullptr, (int64_t*)nullptr, (int32_t*)nullptr, (int)n);
  return align256(a > b ? a : b) + 2 * align256((size_t)n * 8) + 512;
}

int oc_prepare(const int64_t* object_id, const uint8_t* object_mask, int64_t n, int64_t* uniq, int32_t* obj_slot,
               int32_t* n_uniq, void* ws, size_t ws_bytes, cudaStream_t st) {
  GTB_REQUIRE(n >= 1 && n < (1ll << 31) - 1, GTB_ERR_BAD_ARG, "gtb_oc_prepare: bad n_nodes");
  GTB_REQUIRE(ws_bytes >= oc_workspace_bytes(n), GTB_ERR_WORKSPACE, "gtb_oc_prepare: workspace too small");
  char* p = static_cast<char*>(ws);
  int64_t* keys = reinterpret_cast<int64_t*>(p);
  p += align256((size_t)n * 8);
  int64_t* sorted = reinterpret_cast<int64_t*>(p);
  p += align256((size_t)n * 8);
  int32_t* n_sel = reinterpret_cast<int32_t*>(p);
  p += 256;
  size_t cub_bytes = ws_bytes - (p - static_cast<char*>(ws));
  const int threads = 256;
  const int blocks = (int)imin64((n + threads - 1) / threads, (int64_t)kNumSMs * 16);
  oc_keys_kernel<<<blocks, threads, 0, st>>>(object_id, object_mask, n, keys);
  GTB_CHECK_LAUNCH("oc_keys_kernel");
  int rc = check_cuda(cub::DeviceRadixSort::SortKeys(p, cub_bytes, keys, sorted, (int)n, 0, 64, st), "SortKeys");
  if (rc) return rc;
  rc = check_cuda(cub::DeviceSelect::Unique(p, cub_bytes, sorted, uniq, n_sel, (int)n, st), "Unique");
  if (rc) return rc;
  oc_slots_kernel<<<blocks, threads, 0, st>>>(object_id, n, uniq, n_sel, n_uniq, obj_slot);
  GTB_CHECK_LAUNCH("oc_slots_kernel");
  return GTB_OK;
}

__device__ __forceinline__ float oc_charge(float beta, float q_min) {
  const float a = atanhf(beta);
  return a * a + q_min;
}

__global__ void oc_argmax_kernel(const float* __restrict__ beta, const int32_t* __restrict__ slot, int64_t n,
                                 float q_min, unsigned long long* __restrict__ packed) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const int s = slot[j];
    if (s < 0) continue;
    const float q = oc_charge(beta[j], q_min);
    // q > 0: its bit pattern is monotone; ties go to the first index (torch.argmax)
    const unsigned long long key = ((unsigned long long)__float_as_uint(q) << 32) | (0xFFFFFFFFu - (unsigned)j);
    atomicMax(packed + s, key);
  }
}

__global__ void oc_alphas_kernel(const unsigned long long* __restrict__ packed, int k, int32_t* __restrict__ alphas) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k) alphas[i] = (int32_t)(0xFFFFFFFFu - (unsigned)(packed[i] & 0xFFFFFFFFull));
}

int oc_alphas(const float* beta, const int32_t* slot, int64_t n, float q_min, int32_t k,
              unsigned long long* packed, int32_t* alphas, cudaStream_t st) {
  GTB_REQUIRE(k >= 1, GTB_ERR_BAD_ARG, "gtb_oc_alphas: no objects of interest");
  int rc = check_cuda(cudaMemsetAsync(packed, 0, (size_t)k * 8, st), "memset packed");
  if (rc) return rc;
  const int threads = 256;
  const int blocks = (int)imin64((n + threads - 1) / threads, (int64_t)kNumSMs * 16);
  oc_argmax_kernel<<<blocks, threads, 0, st>>>(beta, slot, n, q_min, packed);
  GTB_CHECK_LAUNCH("oc_argmax_kernel");
  oc_alphas_kernel<<<(k + 255) / 256, 256, 0, st>>>(packed, k, alphas);
  GTB_CHECK_LAUNCH("oc_alphas_kernel");
  return GTB_OK;
}

constexpr int OC_TJ = 256;    // hits per block (one per thread)
constexpr int OC_TK = 256;    // condensation points staged per smem tile
constexpr int OC_KSPLIT = 2048;  // condensation points per blockIdx.y

__global__ void __launch_bounds__(OC_TJ)
oc_potentials_kernel(const float* __restrict__ beta, const float* __restrict__ x, int d,
                     const int64_t* __restrict__ object_id, const uint8_t* __restrict__ mask,
                     const int32_t* __restrict__ slot, int64_t n, const int32_t* __restrict__ alphas, int k,
                     float q_min, int64_t noise_thr, double* __restrict__ out) {
  __shared__ float xs[OC_TK * OC_MAXD];
  __shared__ float qs[OC_TK];
  __shared__ double red[32];
  const int64_t j = (int64_t)blockIdx.x * OC_TJ + threadIdx.x;
  const bool valid = j < n;
  float xj[OC_MAXD];
#pragma unroll
  for (int t = 0; t < OC_MAXD; ++t) xj[t] = (valid && t < d) ? x[j * d + t] : 0.f;
  const float qj = valid ? oc_charge(beta[j], q_min) : 0.f;
  const int sj = valid ? slot[j] : -1;
  double v_att = 0.0, v_rep = 0.0, n_rep = 0.0;
  const int k_beg = blockIdx.y * OC_KSPLIT, k_end = min(k, k_beg + OC_KSPLIT);
  for (int kt = k_beg; kt < k_end; kt += OC_TK) {
    const int kn = min(OC_TK, k_end - kt);
    __syncthreads();
    for (int i = threadIdx.x; i < kn; i += OC_TJ) {
      const int a = alphas[kt + i];
      qs[i] = oc_charge(beta[a], q_min);
      for (int t = 0; t < d; ++t) xs[i * d + t] = x[(int64_t)a * d + t];
    }
    __syncthreads();
    if (!valid) continue;
    float att_t = 0.f, rep_t = 0.f;
    int rep_c = 0;
    for (int i = 0; i < kn; ++i) {
      float d2 = 0.f;
      for (int t = 0; t < d; ++t) {
        const float df = xj[t] - xs[i * d + t];
        d2 = fmaf(df, df, d2);
      }
      const float dist = sqrtf(d2);
      const float qq = qj * qs[i];
      if (kt + i == sj) {
        att_t += qq * (dist * dist);
      } else if (dist < 1.f) {
        rep_t += qq * (1.f - dist);
        ++rep_c;
      }
    }
    v_att += att_t; v_rep += rep_t; n_rep += rep_c;
  }
  double coward = 0.0, noise = 0.0, n_noise = 0.0, n_oi = 0.0;
  if (blockIdx.y == 0 && valid) {
    if (j < k) coward = 1.0 - (double)beta[alphas[j]];
    if (!(object_id[j] > noise_thr)) { noise = beta[j]; n_noise = 1.0; }
    if (mask[j]) n_oi = 1.0;
  }
  double vals[7] = {v_att, v_rep, coward, noise, n_noise, n_oi, n_rep};
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    double v = vals[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < OC_TJ / 32; ++w) s += red[w];
      if (s != 0.0) atomicAdd(out + i, s);
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) out[7] = (double)k;
}

int oc_potentials(const float* beta, const float* x, int32_t d, const int64_t* object_id, const uint8_t* mask,
                  const int32_t* slot, int64_t n, const int32_t* alphas, int32_t k, float q_min,
                  int64_t noise_thr, double* out, cudaStream_t st) {
  GTB_REQUIRE(d >= 1 && d <= OC_MAXD, GTB_ERR_UNSUPPORTED_DIM, "gtb_oc_potentials: latent dim %d outside [1,%d]", d,
              OC_MAXD);
  GTB_REQUIRE(k >= 1 && n >= 1, GTB_ERR_BAD_ARG, "gtb_oc_potentials: empty input");
  dim3 grid((unsigned)((n + OC_TJ - 1) / OC_TJ), (unsigned)((k + OC_KSPLIT - 1) / OC_KSPLIT));
  oc_potentials_kernel<<<grid, OC_TJ, 0, st>>>(beta, x, d, object_id, mask, slot, n, alphas, k, q_min, noise_thr, out);
  GTB_CHECK_LAUNCH("oc_potentials_kernel");
  return GTB_OK;
}

// ------------------------------------------------------------------ gradients of the four sums
// What torch autograd derives for oc.py:282-336 (cdist, indexing by alphas_k, the masked sums):
//   attractive pair (k == slot_j):      q_j q_k D^2      -> d/dx_j = 2 q_j q_k (x_j - x_k), d/dq_j = q_k D^2
//   repulsive pair (k != slot_j, D<1):  q_j q_k (1 - D)  -> d/dx_j = -q_j q_k (x_j - x_k)/D (0 at D = 0, as
//                                                           cdist's backward), d/dq_j = q_k (1 - D)
// and the mirror terms for the condensation point.  Hit-side sums: one thread per hit over tiles of
// condensation points; CP-side sums: one thread per condensation point over tiles of hits (the pair
// is evaluated twice rather than reduced across threads with ~N*K atomics).
// coef = {g_att / norm_att, g_rep / norm_rep, g_coward / K, g_noise / n_noise}.
__global__ void __launch_bounds__(OC_TJ)
oc_grad_hits_kernel(const float* __restrict__ beta, const float* __restrict__ x, int d,
                    const int32_t* __restrict__ slot, int64_t n, const int32_t* __restrict__ alphas, int k, float q_min,
                    const float* __restrict__ coef, float* __restrict__ gq, float* __restrict__ gx) {
  __shared__ float xs[OC_TK * OC_MAXD];
  __shared__ float qs[OC_TK];
  const int64_t j = (int64_t)blockIdx.x * OC_TJ + threadIdx.x;
  const bool valid = j < n;
  const float c_att = coef[0], c_rep = coef[1];
  float xj[OC_MAXD], gxj[OC_MAXD];
#pragma unroll
  for (int t = 0; t < OC_MAXD; ++t) {
    xj[t] = (valid && t < d) ? x[j * d + t] : 0.f;
    gxj[t] = 0.f;
  }
  const float qj = valid ? oc_charge(beta[j], q_min) : 0.f;
  const int sj = valid ? slot[j] : -1;
  float gqj = 0.f;
  for (int kt = 0; kt < k; kt += OC_TK) {
    const int kn = min(OC_TK, k - kt);
    __syncthreads();
    for (int i = threadIdx.x; i < kn; i += OC_TJ) {
      const int a = alphas[kt + i];
      qs[i] = oc_charge(beta[a], q_min);
      for (int t = 0; t < d; ++t) xs[i * d + t] = x[(int64_t)a * d + t];
    }
    __syncthreads();
    if (!valid) continue;
    for (int i = 0; i < kn; ++i) {
      float df[OC_MAXD];
      float d2 = 0.f;
#pragma unroll
      for (int t = 0; t < OC_MAXD; ++t) {
        df[t] = t < d ? xj[t] - xs[i * d + t] : 0.f;
        d2 = fmaf(df[t], df[t], d2);
      }
      const float dist = sqrtf(d2);
      float wq, wx;  // d/dq_j = wq, d/dx_j = wx * (x_j - x_k)
      if (kt + i == sj) {
        wq = c_att * qs[i] * (dist * dist);
        wx = 2.f * c_att * qj * qs[i];
      } else if (dist < 1.f) {
        wq = c_rep * qs[i] * (1.f - dist);
        wx = dist > 0.f ? -c_rep * qj * qs[i] / dist : 0.f;
      } else {
        continue;
      }
      gqj += wq;
#pragma unroll
      for (int t = 0; t < OC_MAXD; ++t) gxj[t] = fmaf(wx, df[t], gxj[t]);
    }
  }
  if (!valid) return;
  gq[j] = gqj;
  for (int t = 0; t < d; ++t) gx[j * d + t] = gxj[t];
}

constexpr int OC_JSPLIT = 8192;  // hits per blockIdx.y of the CP-side kernel

__global__ void __launch_bounds__(OC_TK)
oc_grad_cps_kernel(const float* __restrict__ beta, const float* __restrict__ x, int d,
                   const int32_t* __restrict__ slot, int64_t n, const int32_t* __restrict__ alphas, int k, float q_min,
                   const float* __restrict__ coef, float* __restrict__ gq, float* __restrict__ gx) {
  __shared__ float xs[OC_TJ * OC_MAXD];
  __shared__ float qs[OC_TJ];
  __shared__ int ss[OC_TJ];
  const int i = blockIdx.x * OC_TK + threadIdx.x;
  const bool valid = i < k;
  const float c_att = coef[0], c_rep = coef[1];
  const int a = valid ? alphas[i] : 0;
  float xk[OC_MAXD], gxk[OC_MAXD];
#pragma unroll
  for (int t = 0; t < OC_MAXD; ++t) {
    xk[t] = (valid && t < d) ? x[(int64_t)a * d + t] : 0.f;
    gxk[t] = 0.f;
  }
  const float qk = valid ? oc_charge(beta[a], q_min) : 0.f;
  float gqk = 0.f;
  const int64_t j_beg = (int64_t)blockIdx.y * OC_JSPLIT, j_end = min(n, j_beg + OC_JSPLIT);
  for (int64_t jt = j_beg; jt < j_end; jt += OC_TJ) {
    const int jn = (int)min((int64_t)OC_TJ, j_end - jt);
    __syncthreads();
    for (int u = threadIdx.x; u < jn; u += OC_TK) {
      const int64_t j = jt + u;
      qs[u] = oc_charge(beta[j], q_min);
      ss[u] = slot[j];
      for (int t = 0; t < d; ++t) xs[u * d + t] = x[j * d + t];
    }
    __syncthreads();
    if (!valid) continue;
    for (int u = 0; u < jn; ++u) {
      float df[OC_MAXD];
      float d2 = 0.f;
#pragma unroll
      for (int t = 0; t < OC_MAXD; ++t) {
        df[t] = t < d ? xs[u * d + t] - xk[t] : 0.f;  // x_j - x_k
        d2 = fmaf(df[t], df[t], d2);
      }
      const float dist = sqrtf(d2);
      float wq, wx;  // d/dq_k = wq, d/dx_k = -wx * (x_j - x_k)
      if (ss[u] == i) {
        wq = c_att * qs[u] * (dist * dist);
        wx = 2.f * c_att * qs[u] * qk;
      } else if (dist < 1.f) {
        wq = c_rep * qs[u] * (1.f - dist);
        wx = dist > 0.f ? -c_rep * qs[u] * qk / dist : 0.f;
      } else {
        continue;
      }
      gqk += wq;
#pragma unroll
      for (int t = 0; t < OC_MAXD; ++t) gxk[t] = fmaf(-wx, df[t], gxk[t]);
    }
  }
  if (!valid) return;
  atomicAdd(gq + a, gqk);
  for (int t = 0; t < d; ++t) atomicAdd(gx + (int64_t)a * d + t, gxk[t]);
}

// gbeta_j = gq_j dq/dbeta + [noise hit] c_noise - [condensation point] c_coward, q = atanh(beta)^2 + q_min
__global__ void oc_grad_beta_kernel(const float* __restrict__ beta, const int64_t* __restrict__ object_id,
                                    const int32_t* __restrict__ slot, int64_t n, const int32_t* __restrict__ alphas,
                                    int64_t noise_thr, const float* __restrict__ coef, const float* __restrict__ gq,
                                    float* __restrict__ gbeta) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const float b = beta[j];
    float g = gq[j] * 2.f * atanhf(b) / (1.f - b * b);
    if (!(object_id[j] > noise_thr)) g += coef[3];
    const int s = slot[j];
    if (s >= 0 && alphas[s] == (int32_t)j) g -= coef[2];
    gbeta[j] = g;
  }
}

int oc_potentials_grad(const float* beta, const float* x, int32_t d, const int64_t* object_id, const int32_t* slot,
                       int64_t n, const int32_t* alphas, int32_t k, float q_min, int64_t noise_thr, const float* coef,
                       float* gq, float* gbeta, float* gx, cudaStream_t st) {
  GTB_REQUIRE(d >= 1 && d <= OC_MAXD, GTB_ERR_UNSUPPORTED_DIM, "gtb_oc_potentials_grad: latent dim %d outside [1,%d]", d,
              OC_MAXD);
  GTB_REQUIRE(k >= 1 && n >= 1 && coef && gq && gbeta && gx, GTB_ERR_BAD_ARG, "gtb_oc_potentials_grad: bad arguments");
  oc_grad_hits_kernel<<<(unsigned)((n + OC_TJ - 1) / OC_TJ), OC_TJ, 0, st>>>(beta, x, d, slot, n, alphas, k, q_min, coef, gq,
                                                                             gx);
  GTB_CHECK_LAUNCH("oc_grad_hits_kernel");
  dim3 grid((unsigned)((k + OC_TK - 1) / OC_TK), (unsigned)((n + OC_JSPLIT - 1) / OC_JSPLIT));
  oc_grad_cps_kernel<<<grid, OC_TK, 0, st>>>(beta, x, d, slot, n, alphas, k, q_min, coef, gq, gx);
  GTB_CHECK_LAUNCH("oc_grad_cps_kernel");
  const int blocks = (int)imin64((n + 255) / 256, (int64_t)kNumSMs * 16);
  oc_grad_beta_kernel<<<blocks, 256, 0, st>>>(beta, object_id, slot, n, alphas, noise_thr, coef, gq, gbeta);
  GTB_CHECK_LAUNCH("oc_grad_beta_kernel");
  return GTB_OK;
}

}  // namespace gtb
