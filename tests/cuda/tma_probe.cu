// Stand-alone hardware probe (run on the B200 box) of the building blocks of the warp-specialised
// edge kernels (gnn_tracking_b200/csrc/edge_ws.cu):
//   mode 0: TMA tile::gather4 of fp32 rows (box 32 x 1, 128-byte swizzle) -> is the shared-memory image the
//           K-major SWIZZLE_128B tile the UMMA descriptors expect (chunk c of row r at chunk c ^ (r & 7))?
//   mode 1: TMA tile::scatter4 of the same image back to scattered rows, rows outside the table skipped
//   mode 2: gather4 / scatter4 of bf16 rows (box 64 x 1)
//   mode 3: tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), A operand in TMEM (two bf16 per 32-bit column),
//           B operand K-major SWIZZLE_128B in shared memory
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 tests/cuda/tma_probe.cu -o tests/cuda/tma_probe
// Each mode runs in its own process (`tma_probe <mode>`): a faulting variant poisons its context only.
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../gnn_tracking_b200/csrc/tc_common.cuh"
#include "../../gnn_tracking_b200/csrc/tma_common.cuh"

using namespace gtb::tc;
using namespace gtb;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

// one warp: lane i moves rows 4 i .. 4 i + 3 of a 128-row tile.  SUBS sub-tiles of 128 bytes per row.
template <int SUBS, int BOXC>
__global__ void __launch_bounds__(32) gather_scatter_kernel(const __grid_constant__ CUtensorMap in_map,
                                                            const __grid_constant__ CUtensorMap out_map,
                                                            const int32_t* __restrict__ idx_in, const int32_t* __restrict__ idx_out,
                                                            unsigned char* __restrict__ image, int do_store, int* err) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  const int lane = threadIdx.x;
  const uint32_t s0 = smem_u32(smem), b = smem_u32(&bar);
  if (s0 & 1023u) {
    if (lane == 0) *err = 3;
    return;
  }
  if (lane == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncwarp();
  const int4 r = *reinterpret_cast<const int4*>(idx_in + 4 * lane);
  if (lane == 0) tma::mbar_expect_tx(b, SUBS * 128 * 128);
  __syncwarp();
#pragma unroll
  for (int s = 0; s < SUBS; ++s) tma::gather4(s0 + s * 16384 + lane * 512, &in_map, b, s * BOXC, r.x, r.y, r.z, r.w);
  if (!mbar_wait(&bar, 0)) {
    if (lane == 0) *err = 1;
    return;
  }
  for (int i = lane; i < SUBS * 16384 / 16; i += 32)
    reinterpret_cast<uint4*>(image)[i] = reinterpret_cast<const uint4*>(smem)[i];
  if (do_store) {
    const int4 o = *reinterpret_cast<const int4*>(idx_out + 4 * lane);
    fence_proxy_async_smem();
    __syncwarp();
#pragma unroll
    for (int s = 0; s < SUBS; ++s) tma::scatter4(&out_map, s0 + s * 16384 + lane * 512, s * BOXC, o.x, o.y, o.z, o.w);
    tma::bulk_commit();
    tma::bulk_wait_all0();
  }
}

template <typename T>
static int run_gather(int mode) {
  constexpr bool BF = sizeof(T) == 2;
  constexpr int SUBS = BF ? 1 : 2, BOXC = BF ? 64 : 32;
  const int R = 1000, C = 64;
  std::vector<T> tab((size_t)R * C);
  for (int i = 0; i < R * C; ++i) tab[i] = (T)(float)((i % 4093) * (BF ? 1 : 3) + (BF ? 0 : 0.5f));
  std::vector<int32_t> idx_in(128), idx_out(128);
  srand(7);
  for (int i = 0; i < 128; ++i) idx_in[i] = rand() % R;
  // a permutation-like scatter target with two rows outside the table (must be skipped)
  for (int i = 0; i < 128; ++i) idx_out[i] = (i * 7 + 3) % R;
  idx_out[5] = R + 10;
  idx_out[77] = 0x7fffff00;
  T *d_tab, *d_out;
  int32_t *d_in, *d_o;
  unsigned char* d_img;
  int* d_err;
  CK(cudaMalloc(&d_tab, tab.size() * sizeof(T)));
  CK(cudaMalloc(&d_out, tab.size() * sizeof(T)));
  CK(cudaMalloc(&d_in, 512));
  CK(cudaMalloc(&d_o, 512));
  CK(cudaMalloc(&d_img, SUBS * 16384));
  CK(cudaMalloc(&d_err, 4));
  CK(cudaMemcpy(d_tab, tab.data(), tab.size() * sizeof(T), cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0, tab.size() * sizeof(T)));
  CK(cudaMemcpy(d_in, idx_in.data(), 512, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_o, idx_out.data(), 512, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_err, 0, 4));
  CUtensorMap in_map, out_map;
  const CUtensorMapDataType dt = BF ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  if (!tma::make_map_2d(&in_map, d_tab, dt, sizeof(T), R, C, C, BOXC, 1) ||
      !tma::make_map_2d(&out_map, d_out, dt, sizeof(T), R, C, C, BOXC, 1)) {
    printf("mode %d: cuTensorMapEncodeTiled failed\n", mode);
    return 1;
  }
  CK(cudaFuncSetAttribute(gather_scatter_kernel<SUBS, BOXC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SUBS * 16384));
  gather_scatter_kernel<SUBS, BOXC><<<1, 32, SUBS * 16384>>>(in_map, out_map, d_in, d_o, d_img, mode != 0, d_err);
  cudaError_t e = cudaDeviceSynchronize();
  int herr = 0;
  std::vector<unsigned char> img(SUBS * 16384);
  std::vector<T> out(tab.size());
  if (e == cudaSuccess) {
    CK(cudaMemcpy(&herr, d_err, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(img.data(), d_img, img.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out.data(), d_out, out.size() * sizeof(T), cudaMemcpyDeviceToHost));
  }
  // expected image: element (row r, column k): sub-tile k / BOXC, 16-byte chunk ((k % BOXC) * sizeof(T) / 16) ^ (r & 7)
  long bad_img = 0, bad_out = 0;
  for (int r = 0; r < 128 && e == cudaSuccess; ++r)
    for (int k = 0; k < C; ++k) {
      const int sub = k / BOXC, kb = (k % BOXC) * (int)sizeof(T);
      const size_t off = (size_t)sub * 16384 + (r >> 3) * 1024 + (r & 7) * 128 + ((((kb >> 4) ^ (r & 7)) & 7) << 4) + (kb & 15);
      T v;
      memcpy(&v, &img[off], sizeof(T));
      if ((float)v != (float)tab[(size_t)idx_in[r] * C + k]) ++bad_img;
    }
  if (mode != 0 && e == cudaSuccess) {
    std::vector<float> expect(tab.size(), 0.f);
    for (int r = 0; r < 128; ++r)
      if (idx_out[r] >= 0 && idx_out[r] < R)
        for (int k = 0; k < C; ++k) expect[(size_t)idx_out[r] * C + k] = (float)tab[(size_t)idx_in[r] * C + k];
    for (size_t i = 0; i < out.size(); ++i)
      if ((float)out[i] != expect[i]) ++bad_out;
  }
  printf("mode %d (%s): cuda=%s err=%d  image mismatches %ld / %d  scatter mismatches %ld\n", mode, BF ? "bf16" : "fp32",
         cudaGetErrorString(e), herr, bad_img, 128 * C, bad_out);
  return (e != cudaSuccess) || herr || bad_img || bad_out;
}

// ------------------------------------------------------------------ kind::f16, A in TMEM
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc)
      : "memory");
}

template <int K, int N>
__global__ void __launch_bounds__(128) bf16_probe(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                  float* __restrict__ D, int* err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int KT = K / 64, B_TILE = N * 128;  // 64 bf16 = 128 bytes per row of a K tile
  unsigned char* sB = smem;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + KT * B_TILE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K, kb = (k % 64) * 2;
    const uint32_t off = (k / 64) * B_TILE + (n >> 3) * 1024 + (n & 7) * 128 + ((((kb >> 4) ^ (n & 7)) & 7) << 4) + (kb & 15);
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = B[i];
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t d_col = 0, a_col = 128;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c = 0; c < K / 2; c += 8) {  // column j: elements 2 j (low half) and 2 j + 1 (high half) of row tid
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) {
      const __nv_bfloat16 lo = A[tid * K + 2 * (c + j)], hi = A[tid * K + 2 * (c + j) + 1];
      v[j] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
    }
    tmem_st8(tmem + lane_base + a_col + c, v);
  }
  tmem_st_wait();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    bool acc = false;
    for (int ks = 0; ks < K / 16; ++ks) {  // K = 16 per MMA: 8 TMEM columns of A, 32 bytes of a B row
      const uint64_t bd = make_smem_desc_sw128(smem_u32(sB + (ks / 4) * B_TILE) + (ks % 4) * 32);
      mma_bf16_ts(tmem + d_col, tmem + a_col + 8 * ks, bd, idesc, acc);
      acc = true;
    }
    mma_commit(bar);
  }
  const bool ok = mbar_wait(bar, 0);
  if (!ok) atomicExch(err, 1);
  tc_fence_after_sync();
  if (ok) {
    for (int c = 0; c < N; c += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + lane_base + d_col + c, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) D[tid * N + c + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

template <int K, int N>
static int run_bf16() {
  std::vector<__nv_bfloat16> A(128 * K), B(N * K);
  std::vector<float> D(128 * N, -1.f);
  srand(99);
  for (auto& v : A) v = __float2bfloat16((float)rand() / RAND_MAX * 2.f - 1.f);
  for (auto& v : B) v = __float2bfloat16((float)rand() / RAND_MAX * 2.f - 1.f);
  __nv_bfloat16 *dA, *dB;
  float* dD;
  int* dErr;
  CK(cudaMalloc(&dA, A.size() * 2));
  CK(cudaMalloc(&dB, B.size() * 2));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMalloc(&dErr, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dErr, 0, 4));
  const size_t smem = (K / 64) * N * 128 + 64 + 1024;
  CK(cudaFuncSetAttribute(bf16_probe<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bf16_probe<K, N><<<1, 128, smem>>>(dA, dB, dD, dErr);
  cudaError_t e = cudaDeviceSynchronize();
  int herr = 0;
  double emax = 0, scale = 0;
  if (e == cudaSuccess) {
    CK(cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double ex = 0;
        for (int k = 0; k < K; ++k) ex += (double)__bfloat162float(A[m * K + k]) * (double)__bfloat162float(B[n * K + k]);
        emax = fmax(emax, fabs(D[m * N + n] - ex));
        scale = fmax(scale, fabs(ex));
      }
  }
  printf("mode 3 bf16 TS K=%d N=%d: cuda=%s mbar_timeout=%d  max|err| vs fp64 of the bf16 operands %.3e (max|ref| %.3f)\n", K, N,
         cudaGetErrorString(e), herr, emax, scale);
  return (e != cudaSuccess) || herr || !(emax < 1e-4 * fmax(1.0, scale));
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  int bad = 0;
  if (mode == 0 || mode == 1) bad = run_gather<float>(mode);
  else if (mode == 2) bad = run_gather<__nv_bfloat16>(mode);
  else if (mode == 3) bad = run_bf16<64, 64>() | run_bf16<128, 128>() | run_bf16<64, 32>();
  printf(bad ? "TMA_PROBE mode %d FAILED\n" : "TMA_PROBE mode %d OK\n", mode);
  return bad;
}
