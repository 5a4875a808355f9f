// Fused row-MLP on fp32 CUDA cores (FFMA tiles): gather -> concat-free MLP (<= 3 Linear
// layers, ReLU between) -> epilogue (activation / residual / row scatter) -> optional
// in-tile segmented sum by destination.  This is the exact-fp32 implementation used for
// any width <= 128 (reference defaults Dn=5, De=4) and the numerical fallback of the
// tcgen05 path.  See include/gtb200.h for the reference citations of each use.
#include "common.cuh"

namespace gtb {

constexpr int TM = 128;        // rows (edges / nodes) per tile
constexpr int KC = 32;         // K chunk
constexpr int AS = KC + 4;     // row stride of the staged A chunk (floats)
constexpr int NTHREADS = 256;

struct SrcS {
  const float* ptr;
  int off;    // first K column of this block
  int width;
  int ld;
  int relu;
  int sidx;   // position of the block in the descriptor (row index table)
};

template <int NC>
struct Smem {
  static constexpr int NW = 64 * NC;
  static constexpr int HS = NW + 4;
  static constexpr size_t a_floats = (size_t)TM * AS;
  static constexpr size_t w_floats = (size_t)KC * NW;
  static constexpr size_t h_floats = (size_t)TM * HS;
  static size_t bytes(int n_srcs) {
    return (a_floats + w_floats + h_floats) * 4 + (size_t)(n_srcs + 2) * TM * 4 + GTB_MAX_SRCS * sizeof(SrcS) + 16;
  }
};

// acc[i][c*4+j] += A[row ty+16i][k] * W[k][c*64 + tx*4 + j]  over one 32-wide K chunk
template <int NC>
__device__ __forceinline__ void gemm_chunk_wide(const float* __restrict__ A, int lda,
                                                const float* __restrict__ Ws, float (&acc)[8][4 * NC],
                                                int tx, int ty) {
  constexpr int NW = 64 * NC;
#pragma unroll
  for (int kk = 0; kk < KC; kk += 4) {
    float4 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ty + 16 * i) * lda + kk);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float* w = Ws + kk * NW + c * 64 + tx * 4;
      const float4 b0 = *reinterpret_cast<const float4*>(w);
      const float4 b1 = *reinterpret_cast<const float4*>(w + NW);
      const float4 b2 = *reinterpret_cast<const float4*>(w + 2 * NW);
      const float4 b3 = *reinterpret_cast<const float4*>(w + 3 * NW);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float* o = &acc[i][c * 4];
        o[0] = fmaf(a[i].x, b0.x, o[0]); o[1] = fmaf(a[i].x, b0.y, o[1]);
        o[2] = fmaf(a[i].x, b0.z, o[2]); o[3] = fmaf(a[i].x, b0.w, o[3]);
        o[0] = fmaf(a[i].y, b1.x, o[0]); o[1] = fmaf(a[i].y, b1.y, o[1]);
        o[2] = fmaf(a[i].y, b1.z, o[2]); o[3] = fmaf(a[i].y, b1.w, o[3]);
        o[0] = fmaf(a[i].z, b2.x, o[0]); o[1] = fmaf(a[i].z, b2.y, o[1]);
        o[2] = fmaf(a[i].z, b2.z, o[2]); o[3] = fmaf(a[i].z, b2.w, o[3]);
        o[0] = fmaf(a[i].w, b3.x, o[0]); o[1] = fmaf(a[i].w, b3.y, o[1]);
        o[2] = fmaf(a[i].w, b3.z, o[2]); o[3] = fmaf(a[i].w, b3.w, o[3]);
      }
    }
  }
}

// narrow last layer (N <= 8, padded to 8): thread = (row tid&127, column half tid>>7)
__device__ __forceinline__ void gemm_chunk_narrow(const float* __restrict__ A, int lda,
                                                  const float* __restrict__ Ws /*[KC][8]*/, float (&acc)[4],
                                                  int r, int h) {
#pragma unroll
  for (int kk = 0; kk < KC; kk += 4) {
    const float4 a = *reinterpret_cast<const float4*>(A + r * lda + kk);
    const float4 b0 = *reinterpret_cast<const float4*>(Ws + (kk + 0) * 8 + h * 4);
    const float4 b1 = *reinterpret_cast<const float4*>(Ws + (kk + 1) * 8 + h * 4);
    const float4 b2 = *reinterpret_cast<const float4*>(Ws + (kk + 2) * 8 + h * 4);
    const float4 b3 = *reinterpret_cast<const float4*>(Ws + (kk + 3) * 8 + h * 4);
    acc[0] = fmaf(a.x, b0.x, acc[0]); acc[1] = fmaf(a.x, b0.y, acc[1]); acc[2] = fmaf(a.x, b0.z, acc[2]); acc[3] = fmaf(a.x, b0.w, acc[3]);
    acc[0] = fmaf(a.y, b1.x, acc[0]); acc[1] = fmaf(a.y, b1.y, acc[1]); acc[2] = fmaf(a.y, b1.z, acc[2]); acc[3] = fmaf(a.y, b1.w, acc[3]);
    acc[0] = fmaf(a.z, b2.x, acc[0]); acc[1] = fmaf(a.z, b2.y, acc[1]); acc[2] = fmaf(a.z, b2.z, acc[2]); acc[3] = fmaf(a.z, b2.w, acc[3]);
    acc[0] = fmaf(a.w, b3.x, acc[0]); acc[1] = fmaf(a.w, b3.y, acc[1]); acc[2] = fmaf(a.w, b3.z, acc[2]); acc[3] = fmaf(a.w, b3.w, acc[3]);
  }
}

template <int NC>
__global__ void __launch_bounds__(NTHREADS, NC == 1 ? 2 : 1)
fused_mlp_ffma_kernel(const __grid_constant__ gtb_mlp_desc_t d, const __grid_constant__ FfmaLayout L) {
  using S = Smem<NC>;
  constexpr int HS = S::HS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);
  float* Ws = As + S::a_floats;
  float* Hs = Ws + S::w_floats;
  int32_t* ridx = reinterpret_cast<int32_t*>(Hs + S::h_floats);  // [n_srcs][TM]
  int32_t* orow = ridx + d.n_srcs * TM;                          // [TM]
  int32_t* segs = orow + TM;                                     // [TM]
  SrcS* srcs = reinterpret_cast<SrcS*>(segs + TM);               // [n_stream] concatenated (non-projected) blocks
  int* n_stream_p = reinterpret_cast<int*>(srcs + GTB_MAX_SRCS);

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * TM;
  const int rows_here = (int)min((int64_t)TM, d.n_rows - row0);
  const float* __restrict__ packed = static_cast<const float*>(d.packed);

  if (tid == 0) {
    int off = 0, n = 0;
    for (int s = 0; s < d.n_srcs; ++s) {
      if (d.srcs[s].flags & GTB_SRC_PROJECTED) continue;
      srcs[n++] = SrcS{d.srcs[s].ptr, off, d.srcs[s].width, d.srcs[s].ld, d.srcs[s].relu, s};
      off += d.srcs[s].width;
    }
    *n_stream_p = n;
  }
  for (int i = tid; i < d.n_srcs * TM; i += NTHREADS) {
    const int s = i / TM, r = i - s * TM;
    int v = 0;
    if (r < rows_here) v = d.srcs[s].index ? d.srcs[s].index[row0 + r] : (int)(row0 + r);
    ridx[i] = v;
  }
  if (tid < TM) {
    int o = 0, sg = -1;
    if (tid < rows_here) {
      o = d.out_index ? d.out_index[row0 + tid] : (int)(row0 + tid);
      if (d.seg_id) sg = d.seg_id[row0 + tid];
    }
    orow[tid] = o;
    segs[tid] = sg;
  }
  __syncthreads();

  const int last = d.n_layers - 1;
  const int n_stream = *n_stream_p;
  for (int l = 0; l <= last; ++l) {
    const bool narrow = (l == last) && L.narrow_last;
    const int Kp = L.kp[l];
    const float* __restrict__ Wg = packed + L.w_off[l];
    const float* __restrict__ bg = packed + L.b_off[l];
    const int nw = L.nw[l];
    float acc[8][4 * NC];
    float nacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4 * NC; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < Kp; k0 += KC) {
      // ---- stage the weight chunk: KC x nw contiguous floats
      {
        const float4* g = reinterpret_cast<const float4*>(Wg + (size_t)k0 * nw);
        float4* s4 = reinterpret_cast<float4*>(Ws);
        const int n4 = KC * nw / 4;
        for (int i = tid; i < n4; i += NTHREADS) s4[i] = __ldg(g + i);
      }
      // ---- stage the gathered A chunk (layer 0 only): lane = k, warp strides rows
      if (l == 0) {
        const int k = k0 + lane;
        int s = -1;
        if (k < d.dims[0]) {
          s = 0;
          while (s + 1 < n_stream && k >= srcs[s + 1].off) ++s;
        }
        if (s >= 0) {
          const SrcS sd = srcs[s];
          const int col = k - sd.off;
          const int32_t* ri = ridx + sd.sidx * TM;
#pragma unroll 4
          for (int r = warp; r < TM; r += NTHREADS / 32) {
            float v = 0.f;
            if (r < rows_here) {
              v = __ldg(sd.ptr + (size_t)ri[r] * sd.ld + col);
              if (sd.relu) v = fmaxf(v, 0.f);
              if (d.row_scale) v *= __ldg(d.row_scale + row0 + r);
            }
            As[r * AS + lane] = v;
          }
        } else {
          for (int r = warp; r < TM; r += NTHREADS / 32) As[r * AS + lane] = 0.f;
        }
      }
      __syncthreads();
      const float* A = (l == 0) ? As : (Hs + k0);
      const int lda = (l == 0) ? AS : HS;
      if (narrow) gemm_chunk_narrow(A, lda, Ws, nacc, tid & (TM - 1), tid >> 7);
      else        gemm_chunk_wide<NC>(A, lda, Ws, acc, tx, ty);
      __syncthreads();
    }

    // ---- layer epilogue: bias (+ ReLU between layers) -> Hs
    if (narrow) {
      const int r = tid & (TM - 1), h = tid >> 7;
      const float4 b = *reinterpret_cast<const float4*>(bg + h * 4);
      float4 v = make_float4(nacc[0] + b.x, nacc[1] + b.y, nacc[2] + b.z, nacc[3] + b.w);
      *reinterpret_cast<float4*>(Hs + r * HS + h * 4) = v;
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float4 b = *reinterpret_cast<const float4*>(bg + c * 64 + tx * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 v = make_float4(acc[i][c * 4 + 0] + b.x, acc[i][c * 4 + 1] + b.y,
                                 acc[i][c * 4 + 2] + b.z, acc[i][c * 4 + 3] + b.w);
          if (l == 0 && n_stream != d.n_srcs) {  // gathered rows of the pre-projected blocks
            const int col = c * 64 + tx * 4, n0 = d.dims[1];
            for (int s = 0; s < d.n_srcs; ++s) {
              if (!(d.srcs[s].flags & GTB_SRC_PROJECTED)) continue;
              const float* pr = d.srcs[s].ptr + (size_t)ridx[s * TM + ty + 16 * i] * d.srcs[s].ld + col;
              if (col + 0 < n0) v.x += __ldg(pr + 0);
              if (col + 1 < n0) v.y += __ldg(pr + 1);
              if (col + 2 < n0) v.z += __ldg(pr + 2);
              if (col + 3 < n0) v.w += __ldg(pr + 3);
            }
          }
          if (l != last) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
          }
          *reinterpret_cast<float4*>(Hs + (ty + 16 * i) * HS + c * 64 + tx * 4) = v;
        }
      }
    }
    __syncthreads();
  }

  // ---- final pass: activation, residual, (scattered) store
  const int N = d.dims[d.n_layers];
  const bool want_aggr = d.aggr != nullptr;
  for (int idx = tid; idx < rows_here * N; idx += NTHREADS) {
    const int r = idx / N, n = idx - r * N;
    float v = Hs[r * HS + n];
    if (d.final_act == GTB_ACT_RELU) v = fmaxf(v, 0.f);
    else if (d.final_act == GTB_ACT_SIGMOID_AFFINE) v = d.act_eps + (1.f - 2.f * d.act_eps) * (1.f / (1.f + expf(-v)));
    v *= d.res_b;
    if (d.res) v = fmaf(d.res_a, __ldg(d.res + (size_t)(row0 + r) * d.res_ld + n), v);
    if (d.out_scale) v *= __ldg(d.out_scale);
    if (d.gate && !(__ldg(d.gate + (size_t)(row0 + r) * d.gate_ld + n) > 0.f)) v = 0.f;
    if (d.out) d.out[(size_t)orow[r] * d.out_ld + n] = v;
    if (want_aggr) Hs[r * HS + n] = v;
  }
  if (!want_aggr) return;
  __syncthreads();

  // ---- in-tile segmented sum by destination (rows are dst-sorted).  A run that covers a
  // node's whole CSR range is stored; partial runs (tile / part boundaries) are added atomically.
  constexpr int RP = 16;  // rows per part
  const int n_parts = (rows_here + RP - 1) / RP;
  for (int it = tid; it < N * n_parts; it += NTHREADS) {
    const int part = it / N, c = it - part * N;
    const int r_beg = part * RP, r_end = min(r_beg + RP, rows_here);
    int cur = segs[r_beg];
    int g_start = r_beg;
    float sum = 0.f;
    for (int r = r_beg; r <= r_end; ++r) {
      const int sg = (r < r_end) ? segs[r] : -2;
      if (sg != cur) {
        const int64_t gs = row0 + g_start, ge = row0 + r;
        float* dst = d.aggr + (size_t)cur * d.aggr_ld + c;
        if (d.rowptr[cur] == gs && d.rowptr[cur + 1] == ge) *dst = sum;
        else atomicAdd(dst, sum);
        cur = sg; g_start = r; sum = 0.f;
      }
      if (r < r_end) sum += Hs[r * HS + c];
    }
  }
}

__global__ void pack_ffma_kernel(const float* __restrict__ W, const float* __restrict__ b, int K, int N,
                                 int Kp, int Nw, float* __restrict__ Wt, float* __restrict__ bt) {
  const int total = Kp * Nw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total + Nw; i += gridDim.x * blockDim.x) {
    if (i < total) {
      const int k = i / Nw, n = i - k * Nw;
      Wt[i] = (k < K && n < N) ? W[(size_t)n * K + k] : 0.f;
    } else {
      const int n = i - total;
      bt[n] = (b != nullptr && n < N) ? b[n] : 0.f;
    }
  }
}

int pack_ffma(int n_layers, const int32_t* dims, const float* const* weights, const float* const* biases,
              void* packed, cudaStream_t st) {
  FfmaLayout L;
  GTB_REQUIRE(ffma_layout(n_layers, dims, &L), GTB_ERR_UNSUPPORTED_DIM,
              "gtb_mlp_pack: Linear widths above %d are not supported by the FFMA path", GTB_MAX_WIDTH);
  float* p = static_cast<float*>(packed);
  for (int l = 0; l < n_layers; ++l) {
    // for l > 0 the K extent of the packed matrix is the padded previous width; rows >= dims[l] are zero
    const int total = L.kp[l] * L.nw[l] + L.nw[l];
    pack_ffma_kernel<<<(total + 255) / 256, 256, 0, st>>>(weights[l], biases ? biases[l] : nullptr, dims[l],
                                                         dims[l + 1], L.kp[l], L.nw[l], p + L.w_off[l],
                                                         p + L.b_off[l]);
    GTB_CHECK_LAUNCH("gtb_mlp_pack");
  }
  return GTB_OK;
}

template <int NC>
static int launch_ffma(const gtb_mlp_desc_t& d, const FfmaLayout& L, cudaStream_t st) {
  const size_t smem = Smem<NC>::bytes(d.n_srcs);
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fused_mlp_ffma_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Smem<NC>::bytes(GTB_MAX_SRCS));
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(fused_mlp_ffma)");
    configured = true;
  }
  const int64_t tiles = (d.n_rows + TM - 1) / TM;
  fused_mlp_ffma_kernel<NC><<<(unsigned)tiles, NTHREADS, smem, st>>>(d, L);
  GTB_CHECK_LAUNCH("fused_mlp_ffma_kernel");
  return GTB_OK;
}

int fused_mlp_ffma(const gtb_mlp_desc_t& d, cudaStream_t st) {
  FfmaLayout L;
  GTB_REQUIRE(ffma_layout(d.n_layers, d.dims, &L), GTB_ERR_UNSUPPORTED_DIM,
              "gtb_fused_mlp_f32: Linear widths above %d are not supported by the FFMA path", GTB_MAX_WIDTH);
  if (d.n_rows == 0) return GTB_OK;
  return L.nc == 1 ? launch_ffma<1>(d, L, st) : launch_ffma<2>(d, L, st);
}

}  // namespace gtb
