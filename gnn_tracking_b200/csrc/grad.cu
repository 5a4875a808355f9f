// Building blocks of the backward pass of the fused row MLP (recompute-based, see
// gnn_tracking_b200/autograd.py).  The activation-gradient products dX = dY W reuse the forward
// tiles (gtb_fused_mlp_f32 with transposed weights and the `gate` epilogue); this file holds what
// has no forward counterpart:
//   * gtb_rows_atb_f32        : weight / bias gradients, out[Ka, Nb] += sum_r act(A[ia(r)])^T B[r]
//   * gtb_rows_scatter_add_f32: gradient of a row gather, dst[index[r]] += src[r]
// Reference semantics: torch autograd through models/mlp.py:59-62 and the index_select /
// torch.cat / scatter_add_ chain of models/interaction_network.py:67-103.
#include "common.cuh"

namespace gtb {

int rows_atb_tc(const float*, int, const int32_t*, int, int, const float*, int, int, int64_t, float*, int, float*, cudaStream_t, bool*);

constexpr int ATB_ROWS = 64;     // rows per staged tile
constexpr int ATB_THREADS = 256;

typedef unsigned long long atb_f32x2;
__device__ __forceinline__ atb_f32x2 atb_pack2(float a, float b) {
  atb_f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ atb_f32x2 atb_fma2(atb_f32x2 a, atb_f32x2 b, atb_f32x2 c) {
  atb_f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// A tile [64][KA] and B tile [64][NB] in shared memory (KA, NB <= 64, zero padded to 64).  Two
// groups of 128 threads take the even / odd rows of a tile; thread (ki = t >> 4, nj = t & 15) of a
// group owns the 8 x 4 block out[8 ki .., 4 nj ..] as packed pairs: per row 3 x LDS.128 and
// 16 x FFMA2 (a_i broadcast against the pairs (b0, b1), (b2, b3); ptxas folds the broadcast into the
// instruction's scalar operand), i.e. 19 issue slots for 32 multiply-adds.
__global__ void __launch_bounds__(ATB_THREADS, 3) rows_atb_kernel(const float* __restrict__ A, int a_ld,
                                                               const int32_t* __restrict__ a_index, int a_relu, int ka,
                                                               const float* __restrict__ B, int b_ld, int nb,
                                                               int64_t n_rows, float* __restrict__ out, int out_ld,
                                                               float* __restrict__ colsum) {
  __shared__ __align__(16) float As[ATB_ROWS][68];
  __shared__ __align__(16) float Bs[ATB_ROWS][68];
  const int tid = threadIdx.x, grp = tid >> 7, ki = (tid & 127) >> 4, nj = tid & 15;
  atb_f32x2 acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = atb_pack2(0.f, 0.f);
  float csum = 0.f;  // column tid of B (tid < nb)
  const bool vec_a = (ka & 3) == 0 && (a_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
  const bool vec_b = (nb & 3) == 0 && (b_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0;
  const int64_t n_tiles = (n_rows + ATB_ROWS - 1) / ATB_ROWS;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * ATB_ROWS;
    const int rows_here = (int)min((int64_t)ATB_ROWS, n_rows - row0);
    if (vec_a) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = tid + j * ATB_THREADS, r = i >> 4, c = (i & 15) << 2;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows_here && c < ka) {
          const int64_t ar = a_index ? (int64_t)__ldg(a_index + row0 + r) : row0 + r;
          a = __ldg(reinterpret_cast<const float4*>(A + (size_t)ar * a_ld + c));
          if (a_relu) a = make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
        }
        *reinterpret_cast<float4*>(&As[r][c]) = a;
      }
    } else {
      for (int i = tid; i < ATB_ROWS * 64; i += ATB_THREADS) {
        const int r = i >> 6, c = i & 63;
        float a = 0.f;
        if (r < rows_here && c < ka) {
          const int64_t ar = a_index ? (int64_t)__ldg(a_index + row0 + r) : row0 + r;
          a = __ldg(A + (size_t)ar * a_ld + c);
          if (a_relu) a = fmaxf(a, 0.f);
        }
        As[r][c] = a;
      }
    }
    if (vec_b) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = tid + j * ATB_THREADS, r = i >> 4, c = (i & 15) << 2;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows_here && c < nb) b = __ldg(reinterpret_cast<const float4*>(B + (size_t)(row0 + r) * b_ld + c));
        *reinterpret_cast<float4*>(&Bs[r][c]) = b;
      }
    } else {
      for (int i = tid; i < ATB_ROWS * 64; i += ATB_THREADS) {
        const int r = i >> 6, c = i & 63;
        Bs[r][c] = (r < rows_here && c < nb) ? __ldg(B + (size_t)(row0 + r) * b_ld + c) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int r = grp; r < ATB_ROWS; r += 2) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[r][8 * ki]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[r][8 * ki + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[r][4 * nj]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const atb_f32x2 b01 = atb_pack2(b.x, b.y), b23 = atb_pack2(b.z, b.w);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const atb_f32x2 aa = atb_pack2(av[i], av[i]);
        acc[i][0] = atb_fma2(aa, b01, acc[i][0]);
        acc[i][1] = atb_fma2(aa, b23, acc[i][1]);
      }
    }
    if (colsum != nullptr && tid < nb)
      for (int r = 0; r < rows_here; ++r) csum += Bs[r][tid];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v0, v1;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(acc[i][h]));
      const int k = 8 * ki + i, n = 4 * nj + 2 * h;
      if (k < ka && n < nb) atomicAdd(out + (size_t)k * out_ld + n, v0);
      if (k < ka && n + 1 < nb) atomicAdd(out + (size_t)k * out_ld + n + 1, v1);
    }
  if (colsum != nullptr && tid < nb) atomicAdd(colsum + tid, csum);
}

// One side at most 4 columns wide (the K = 4 first Linear of the edge encoder, the 64 -> 1 last Linear of the W head):
// a matrix-vector-like product that the 64 x 64 tiles above pad sixteen-fold.  Thread = one column of the WIDE side
// with <= 4 accumulators, four row groups per block; the narrow side's row is a broadcast load.  Memory-bound.
template <bool NARROW_A>
__global__ void __launch_bounds__(256) rows_atb_narrow_kernel(const float* __restrict__ A, int a_ld, const int32_t* __restrict__ a_index,
                                                              int a_relu, int ka, const float* __restrict__ B, int b_ld, int nb,
                                                              int64_t n_rows, float* __restrict__ out, int out_ld,
                                                              float* __restrict__ colsum) {
  __shared__ float red[4][64][4];
  const int c = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int wide = NARROW_A ? nb : ka, narrow = NARROW_A ? ka : nb;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, cs[4] = {0.f, 0.f, 0.f, 0.f};
  // four rows per trip, their loads issued together (the trip is latency-bound otherwise: one 4-byte load per thread)
  const int64_t step = (int64_t)gridDim.x * 4;
  for (int64_t r0 = (int64_t)blockIdx.x * 4 + g; r0 < n_rows; r0 += 4 * step) {
    float wv[4], nv[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = r0 + u * step;
      const bool ok = r < n_rows;
      const int64_t rc = ok ? r : n_rows - 1;  // clamped: no branch between the loads; the contribution is zeroed below
      const int64_t ar = a_index ? (int64_t)__ldg(a_index + rc) : rc;
      if (NARROW_A) {
        wv[u] = (ok && c < wide) ? __ldg(B + (size_t)rc * b_ld + c) : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) nv[u][k] = k < narrow ? __ldg(A + (size_t)ar * a_ld + k) : 0.f;
      } else {
        wv[u] = (ok && c < wide) ? __ldg(A + (size_t)ar * a_ld + c) : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) nv[u][k] = (ok && k < narrow) ? __ldg(B + (size_t)rc * b_ld + k) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (NARROW_A) {
        cs[0] += wv[u];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = fmaf(a_relu ? fmaxf(nv[u][k], 0.f) : nv[u][k], wv[u], acc[k]);
      } else {
        const float a = a_relu ? fmaxf(wv[u], 0.f) : wv[u];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc[k] = fmaf(a, nv[u][k], acc[k]);
          cs[k] += nv[u][k];
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) red[g][c][k] = acc[k];
  __syncthreads();
  if (g == 0 && c < wide) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < narrow) {
        const float v = red[0][c][k] + red[1][c][k] + red[2][c][k] + red[3][c][k];
        if (NARROW_A) atomicAdd(out + (size_t)k * out_ld + c, v);
        else          atomicAdd(out + (size_t)c * out_ld + k, v);
      }
    }
  }
  if (colsum != nullptr) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) red[g][c][k] = cs[k];
    __syncthreads();
    if (NARROW_A) {
      if (g == 0 && c < wide) atomicAdd(colsum + c, red[0][c][0] + red[1][c][0] + red[2][c][0] + red[3][c][0]);
    } else if (g == 0 && c == 0) {  // every thread of a row group summed the same b values: take column 0's
      for (int j = 0; j < narrow; ++j) atomicAdd(colsum + j, red[0][0][j] + red[1][0][j] + red[2][0][j] + red[3][0][j]);
    }
  }
}

int rows_atb(const float* A, int a_ld, const int32_t* a_index, int a_relu, int ka, const float* B, int b_ld, int nb,
             int64_t n_rows, float* out, int out_ld, float* colsum, cudaStream_t st) {
  GTB_REQUIRE(A && B && out && ka >= 1 && ka <= 64 && nb >= 1 && nb <= 64 && a_ld >= ka && b_ld >= nb && out_ld >= nb,
              GTB_ERR_BAD_ARG, "gtb_rows_atb_f32: widths must be in [1, 64] (got %d x %d)", ka, nb);
  if (n_rows == 0) return GTB_OK;
  {  // 64 x 64 blocks over many rows: the tensor-core kernel (atb_tc.cu)
    bool handled = false;
    const int rc = rows_atb_tc(A, a_ld, a_index, a_relu, ka, B, b_ld, nb, n_rows, out, out_ld, colsum, st, &handled);
    if (rc != GTB_OK || handled) return rc;
  }
  if ((ka <= 4 || nb <= 4) && n_rows >= 1024) {
    const int grid = (int)imin64((n_rows + 63) / 64, (int64_t)kNumSMs * 8);
    if (ka <= 4) rows_atb_narrow_kernel<true><<<grid, 256, 0, st>>>(A, a_ld, a_index, a_relu, ka, B, b_ld, nb, n_rows, out, out_ld, colsum);
    else         rows_atb_narrow_kernel<false><<<grid, 256, 0, st>>>(A, a_ld, a_index, a_relu, ka, B, b_ld, nb, n_rows, out, out_ld, colsum);
    GTB_CHECK_LAUNCH("rows_atb_narrow_kernel");
    return GTB_OK;
  }
  const int64_t n_tiles = (n_rows + ATB_ROWS - 1) / ATB_ROWS;
  // every CTA ends with 64 x 64 atomics onto the same addresses: at least 16 tiles per CTA, and no
  // more CTAs than are resident at once (3 per SM: __launch_bounds__(256, 3) -> 79 registers, no spills)
  const int grid = (int)imin64((n_tiles + 15) / 16, (int64_t)kNumSMs * 3);
  rows_atb_kernel<<<grid, ATB_THREADS, 0, st>>>(A, a_ld, a_index, a_relu, ka, B, b_ld, nb, n_rows, out, out_ld, colsum);
  GTB_CHECK_LAUNCH("rows_atb_kernel");
  return GTB_OK;
}

__global__ void rows_scatter_add_kernel(const float* __restrict__ src, int src_ld, const int32_t* __restrict__ index,
                                        int64_t n_rows, int width, float* __restrict__ dst, int dst_ld) {
  const int64_t total = n_rows * width;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / width;
    const int c = (int)(i - r * width);
    atomicAdd(dst + (size_t)__ldg(index + r) * dst_ld + c, __ldg(src + (size_t)r * src_ld + c));
  }
}

// 16-byte pieces: one vector reduction (red.global.add.v4.f32) per float4 of a row
__global__ void rows_scatter_add_v4_kernel(const float* __restrict__ src, int src_ld, const int32_t* __restrict__ index,
                                           int64_t n_rows, int w4, float* __restrict__ dst, int dst_ld) {
  const int64_t total = n_rows * w4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / w4;
    const int c = (int)(i - r * w4) << 2;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * src_ld + c));
    float* d = dst + (size_t)__ldg(index + r) * dst_ld + c;
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(d), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
}

int rows_scatter_add(const float* src, int src_ld, const int32_t* index, int64_t n_rows, int width, float* dst,
                     int dst_ld, cudaStream_t st) {
  GTB_REQUIRE(src && index && dst && width >= 1, GTB_ERR_BAD_ARG, "gtb_rows_scatter_add_f32: bad arguments");
  if (n_rows == 0) return GTB_OK;
  if ((width & 3) == 0 && (src_ld & 3) == 0 && (dst_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const int64_t total4 = n_rows * (width >> 2);
    const int blocks4 = (int)imin64((total4 + 255) / 256, (int64_t)kNumSMs * 16);
    rows_scatter_add_v4_kernel<<<blocks4, 256, 0, st>>>(src, src_ld, index, n_rows, width >> 2, dst, dst_ld);
    GTB_CHECK_LAUNCH("rows_scatter_add_v4_kernel");
    return GTB_OK;
  }
  const int64_t total = n_rows * width;
  const int blocks = (int)imin64((total + 255) / 256, (int64_t)kNumSMs * 32);
  rows_scatter_add_kernel<<<blocks, 256, 0, st>>>(src, src_ld, index, n_rows, width, dst, dst_ld);
  GTB_CHECK_LAUNCH("rows_scatter_add_kernel");
  return GTB_OK;
}

}  // namespace gtb
