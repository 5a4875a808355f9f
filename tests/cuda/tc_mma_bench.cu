// Micro-benchmark of tcgen05.mma issue / completion cost on one SM (run on the B200 box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 tests/cuda/tc_mma_bench.cu -o tc_mma_bench
// For kind::tf32, M = 128: how long does a chain of `count` MMAs take (a) to issue, (b) to
// complete, as a function of N, of the A operand source (shared memory vs TMEM) and of the
// number of independent accumulators the chain rotates over?
#include <stdio.h>
#include <stdlib.h>

#include "../../gnn_tracking_b200/csrc/tc_common.cuh"

using namespace gtb::tc;

__global__ void __launch_bounds__(128) mma_bench(int n, int ts_mode, int n_acc, int count, int k_advance,
                                                 long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;              // 2 K-tiles [128][32] fp32 = 32 KB
  unsigned char* sB = sA + 32768;        // 2 K-tiles [n][32] fp32 <= 64 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (32768 + 65536) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
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
  {  // zero the A region (columns 384..447) and the accumulators
    uint32_t z[16];
    for (int j = 0; j < 16; ++j) z[j] = 0;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < 512; c += 16) tmem_st16(tmem + lane_base + c, z);
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc_tf32(128, n);
    const uint64_t ad0 = make_smem_desc_sw128(smem_u32(sA));
    const uint64_t bd0 = make_smem_desc_sw128(smem_u32(sB));
    const long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const int ks = k_advance ? (i & 7) : 0;
      const uint32_t d = tmem + (uint32_t)((i % n_acc) * n);
      const uint64_t bd = bd0 + (uint64_t)((ks >> 2) * (n * 128 / 16) + (ks & 3) * 2);
      if (ts_mode) mma_tf32_ts(d, tmem + 384 + 8 * ks, bd, idesc, i >= n_acc);
      else mma_tf32_ss(d, ad0 + (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 2), bd, idesc, i >= n_acc);
    }
    const long long t1 = clock64();
    mma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  const size_t smem = 32768 + 65536 + 64 + 1024;
  cudaFuncSetAttribute(mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int count = 96;
  printf("kind::tf32 M=128, %d MMAs (K = 8 each); cycles: issue loop / until commit arrives; per MMA\n", count);
  for (int ts = 0; ts <= 1; ++ts)
    for (int n : {32, 64, 128, 256})
      for (int n_acc : {1, 2, 3, 4}) {
        if (n * n_acc > 384) continue;
        for (int adv = 0; adv <= 1; ++adv) {
          long long h[2] = {0, 0};
          for (int rep = 0; rep < 2; ++rep) {
            mma_bench<<<1, 128, smem>>>(n, ts, n_acc, count, adv, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
              printf("CUDA error: %s\n", cudaGetErrorString(e));
              return 1;
            }
            cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
          }
          printf("A=%s N=%3d accumulators=%d k_advance=%d : issue %6lld  total %6lld  | per MMA %.1f (math floor %.0f)\n",
                 ts ? "tmem" : "smem", n, n_acc, adv, h[0], h[1], (double)h[1] / count, 128.0 * n / 256.0);
        }
      }
  return 0;
}
