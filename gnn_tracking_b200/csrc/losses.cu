// Edge-classification losses (reference metrics/losses/ec.py) as one fused reduction:
// optional label falsification by pt[edge_index[0]] (ec.py:71-92), per-edge BCE / focal term,
// block reduction, one double atomicAdd per block.
#include "common.cuh"

namespace gtb {

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;
}

__global__ void ec_loss_kernel(const float* __restrict__ w, const void* __restrict__ y, int label_kind,
                               int64_t n_edges, const int64_t* __restrict__ src, const float* __restrict__ pt,
                               float pt_thld, int mode, float alpha, float gamma, float pos_weight,
                               double* __restrict__ out) {
  __shared__ double sh[32];
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += stride) {
    const float we = w[e];
    float yraw = label_kind == 0 ? static_cast<const float*>(y)[e]
                                 : (static_cast<const uint8_t*>(y)[e] ? 1.f : 0.f);
    float yf = yraw;  // falsified label
    if (pt != nullptr) yf = (yraw != 0.f && pt[src[e]] > pt_thld) ? 1.f : 0.f;
    float term;
    if (mode == 0) {
      // F.binary_cross_entropy clamps each log at -100
      const float lw = fmaxf(logf(we), -100.f), l1w = fmaxf(logf(1.f - we), -100.f);
      term = -(yf * lw + (1.f - yf) * l1w);
    } else {
      // _binary_focal_loss ec.py:12-29
      const float target = (mode == 2) ? (yraw != 0.f ? 1.f : 0.f) : yf;
      const float pw = (mode == 2) ? yf : pos_weight;
      const float pn = 1.f - we;
      const float pos = -alpha * pw * powf(pn, gamma) * target * logf(we);
      const float neg = -(1.f - alpha) * powf(we, gamma) * (1.f - target) * logf(pn);
      term = pos + neg;
    }
    acc += (double)term;
  }
  const double tot = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    atomicAdd(out, tot);
    if (blockIdx.x == 0) out[1] = (double)n_edges;
  }
}

// d(mean loss)/dw per edge, times `scale` (= upstream gradient / n_edges): the analytic derivative
// of the terms above (torch's BCE backward: (w - y) / max(w (1 - w), 1e-12)).
__global__ void ec_loss_grad_kernel(const float* __restrict__ w, const void* __restrict__ y, int label_kind,
                                    int64_t n_edges, const int64_t* __restrict__ src, const float* __restrict__ pt,
                                    float pt_thld, int mode, float alpha, float gamma, float pos_weight,
                                    const float* __restrict__ scale, float* __restrict__ dw) {
  const float sc = __ldg(scale);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += stride) {
    const float we = w[e];
    float yraw = label_kind == 0 ? static_cast<const float*>(y)[e]
                                 : (static_cast<const uint8_t*>(y)[e] ? 1.f : 0.f);
    float yf = yraw;
    if (pt != nullptr) yf = (yraw != 0.f && pt[src[e]] > pt_thld) ? 1.f : 0.f;
    float g;
    if (mode == 0) {
      g = (we - yf) / fmaxf(we * (1.f - we), 1e-12f);
    } else {
      const float target = (mode == 2) ? (yraw != 0.f ? 1.f : 0.f) : yf;
      const float pw = (mode == 2) ? yf : pos_weight;
      const float pn = 1.f - we;
      const float dpos = -alpha * pw * target * (-gamma * powf(pn, gamma - 1.f) * logf(we) + powf(pn, gamma) / we);
      const float dneg = -(1.f - alpha) * (1.f - target) * (gamma * powf(we, gamma - 1.f) * logf(pn) - powf(we, gamma) / pn);
      g = dpos + dneg;
    }
    dw[e] = g * sc;
  }
}

int ec_loss_grad(const float* w, const void* y, int label_kind, int64_t n_edges, const int64_t* src, const float* pt,
                 float pt_thld, int mode, float alpha, float gamma, float pos_weight, const float* scale, float* dw,
                 cudaStream_t st) {
  GTB_REQUIRE(mode >= 0 && mode <= 2 && (label_kind == 0 || label_kind == 1) && scale && dw, GTB_ERR_BAD_ARG,
              "gtb_ec_loss_grad_f32: bad arguments");
  if (n_edges == 0) return GTB_OK;
  const int threads = 256;
  const int blocks = (int)imax64(1, imin64((n_edges + threads - 1) / threads, (int64_t)kNumSMs * 8));
  ec_loss_grad_kernel<<<blocks, threads, 0, st>>>(w, y, label_kind, n_edges, src, pt, pt_thld, mode, alpha, gamma,
                                                  pos_weight, scale, dw);
  GTB_CHECK_LAUNCH("ec_loss_grad_kernel");
  return GTB_OK;
}

int ec_loss(const float* w, const void* y, int label_kind, int64_t n_edges, const int64_t* src, const float* pt,
            float pt_thld, int mode, float alpha, float gamma, float pos_weight, double* out, cudaStream_t st) {
  GTB_REQUIRE(mode >= 0 && mode <= 2 && (label_kind == 0 || label_kind == 1), GTB_ERR_BAD_ARG,
              "gtb_ec_loss_f32: bad mode / label_kind");
  GTB_REQUIRE(pt == nullptr || src != nullptr, GTB_ERR_BAD_ARG, "gtb_ec_loss_f32: pt given without edge sources");
  const int threads = 256;
  const int blocks = (int)imax64(1, imin64((n_edges + threads - 1) / threads, (int64_t)kNumSMs * 8));
  ec_loss_kernel<<<blocks, threads, 0, st>>>(w, y, label_kind, n_edges, src, pt, pt_thld, mode, alpha, gamma,
                                             pos_weight, out);
  GTB_CHECK_LAUNCH("ec_loss_kernel");
  return GTB_OK;
}

}  // namespace gtb
