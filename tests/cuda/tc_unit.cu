// Stand-alone hardware probe of the tcgen05 building blocks used by mlp_tc.cu (run on the
// B200 box: `nvcc -gencode arch=compute_100a,code=sm_100a tests/cuda/tc_unit.cu -o tc_unit`).
//   mode 0: SS, raw fp32 operands (answers: does the tensor core truncate or round to tf32?)
//   mode 1: SS, 3xTF32 (hi/lo split of A and B, small terms first)
//   mode 2: TS, 3xTF32 with the A operand written to TMEM by tcgen05.st (lane = row)
// Each mode prints the max error against a float64 reference and exact tf32 models.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../gnn_tracking_b200/csrc/tc_common.cuh"

using namespace gtb::tc;

constexpr int M = 128;

template <int K, int N>
__global__ void __launch_bounds__(128) tc_gemm_probe(const float* __restrict__ A, const float* __restrict__ B,
                                                     float* __restrict__ D, int mode, int* __restrict__ err) {
  constexpr int KT = K / 32;                 // 32-wide K tiles
  constexpr int A_TILE = M * 128, B_TILE = N * 128;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzle atoms need 1024 B alignment
  unsigned char* sA_hi = smem;
  unsigned char* sA_lo = sA_hi + KT * A_TILE;
  unsigned char* sB_hi = sA_lo + KT * A_TILE;
  unsigned char* sB_lo = sB_hi + KT * B_TILE;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB_lo + KT * B_TILE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < M * K; i += 128) {
    const int r = i / K, k = i % K;
    float hi, lo;
    const float v = A[i];
    if (mode == 0) { hi = v; lo = 0.f; } else split_tf32(v, hi, lo);
    const uint32_t off = (k / 32) * A_TILE + sw128_offset(r, k % 32);
    *reinterpret_cast<float*>(sA_hi + off) = hi;
    *reinterpret_cast<float*>(sA_lo + off) = lo;
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    float hi, lo;
    const float v = B[i];
    if (mode == 0) { hi = v; lo = 0.f; } else split_tf32(v, hi, lo);
    const uint32_t off = (k / 32) * B_TILE + sw128_offset(r, k % 32);
    *reinterpret_cast<float*>(sB_hi + off) = hi;
    *reinterpret_cast<float*>(sB_lo + off) = lo;
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t d_col = 0, ahi_col = 128, alo_col = 256;  // column offsets inside the allocation
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  if (mode == 2) {  // A operand into TMEM: thread t owns row t
    for (int c = 0; c < K; c += 16) {
      uint32_t h[16], l[16];
      for (int j = 0; j < 16; ++j) {
        float hi, lo;
        split_tf32(A[tid * K + c + j], hi, lo);
        h[j] = __float_as_uint(hi);
        l[j] = __float_as_uint(lo);
      }
      tmem_st16(tmem + lane_base + ahi_col + c, h);
      tmem_st16(tmem + lane_base + alo_col + c, l);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }

  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(M, N);
    const int n_pass = (mode == 0) ? 1 : 3;
    bool acc = false;
    for (int pass = 0; pass < n_pass; ++pass) {
      // small terms first: lo*hi, hi*lo, then hi*hi
      const bool a_lo = (n_pass == 3 && pass == 0);
      const bool b_lo = (n_pass == 3 && pass == 1);
      for (int kt = 0; kt < KT; ++kt) {
        for (int s = 0; s < 4; ++s) {  // four K = 8 steps per 32-wide tile: +32 bytes each
          const uint64_t bd = make_smem_desc_sw128(smem_u32((b_lo ? sB_lo : sB_hi) + kt * B_TILE) + s * 32);
          if (mode == 2) {
            mma_tf32_ts(tmem + d_col, tmem + (a_lo ? alo_col : ahi_col) + kt * 32 + s * 8, bd, idesc, acc);
          } else {
            const uint64_t ad = make_smem_desc_sw128(smem_u32((a_lo ? sA_lo : sA_hi) + kt * A_TILE) + s * 32);
            mma_tf32_ss(tmem + d_col, ad, bd, idesc, acc);
          }
          acc = true;
        }
      }
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
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static float trunc_tf32(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u &= 0xFFFFE000u;
  memcpy(&v, &u, 4);
  return v;
}
static float rna_tf32(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u += 0x1000u;
  u &= 0xFFFFE000u;
  memcpy(&v, &u, 4);
  return v;
}

template <int K, int N>
int run(int mode) {
  std::vector<float> A(M * K), B(N * K), D(M * N, -1.f);
  srand(1234 + mode);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  int* dErr;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dErr, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, D.size() * 4);
  cudaMemset(dErr, 0, 4);
  const size_t smem = 2 * (K / 32) * (M * 128 + N * 128) + 64 + 1024;
  cudaFuncSetAttribute(tc_gemm_probe<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_gemm_probe<K, N><<<1, 128, smem>>>(dA, dB, dD, mode, dErr);
  cudaError_t e = cudaDeviceSynchronize();
  int herr = 0;
  cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double e_exact = 0, e_trunc = 0, e_rna = 0, scale = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ex = 0, tr = 0, rn = 0;
      for (int k = 0; k < K; ++k) {
        ex += (double)A[m * K + k] * B[n * K + k];
        tr += (double)trunc_tf32(A[m * K + k]) * trunc_tf32(B[n * K + k]);
        rn += (double)rna_tf32(A[m * K + k]) * rna_tf32(B[n * K + k]);
      }
      const double g = D[m * N + n];
      e_exact = fmax(e_exact, fabs(g - ex));
      e_trunc = fmax(e_trunc, fabs(g - tr));
      e_rna = fmax(e_rna, fabs(g - rn));
      scale = fmax(scale, fabs(ex));
    }
  printf("mode %d K=%d N=%d: cuda=%s mbar_timeout=%d  max|err| vs fp64 %.3e  vs trunc-tf32 model %.3e  vs rna-tf32 model %.3e  (max|ref| %.3f)\n",
         mode, K, N, cudaGetErrorString(e), herr, e_exact, e_trunc, e_rna, scale);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dErr);
  return (e != cudaSuccess) || herr;
}

int main() {
  int bad = 0;
  bad |= run<64, 64>(0);
  bad |= run<64, 64>(1);
  bad |= run<64, 64>(2);
  bad |= run<32, 16>(1);
  bad |= run<128, 64>(1);
  bad |= run<64, 128>(2);
  bad |= run<64, 32>(2);
  printf(bad ? "TC_UNIT FAILED\n" : "TC_UNIT DONE\n");
  return bad;
}
