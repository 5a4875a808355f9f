// Weight gradient of one 64 x 64 Linear on the tensor cores (sm_100a):
//
//   out[64, 64] += sum_r act(A[ia(r)])^T B[r]        (and colsum[64] += sum_r B[r]: the bias gradient)
//
// i.e. what torch autograd derives for `nn.Linear.weight` / `.bias` (reference models/mlp.py:59-62) over the
// rows of a fused launch: A = the Linear's input rows (recomputed activations or a gathered feature block),
// B = the gradient of its output rows.  A 2 * 64 * 64 * E flop contraction over E = 1M rows: 512 bytes per row,
// i.e. HBM-bound (~80 us at the copy peak) -- the CUDA-core version of grad.cu needs ~300 us for its 8 GFLOP.
//
// The row index is the K dimension of the product, so both operands are "MN-major" for the tensor core: a
// row-major [rows][32 columns] tile -- what a 2-D TMA tile load leaves in shared memory -- is a canonical
// MN-major operand (32 elements along M / N contiguous, rows along K; cute/atom/mma_traits_sm100.hpp,
// make_umma_desc<Major::MN>; swizzle mode: see at_desc).  No transposes.
//
//   * one persistent CTA per SM, 64-row stages, 3 stages of 64 KB: A | A lo | B | B lo (two 32-column K tiles each);
//   * warp 0: TMA producer (tile loads; tile::gather4 for an indexed A), warp 1: MMA issue, warps 2-9: converters --
//     tf32 hi / lo split of both tiles in place (the tensor core truncates fp32 operands: 3xTF32 keeps the
//     gradient at fp32 level), ReLU of A on the way, column sums of B in registers;
//   * 24 MMAs (M = 128, N = 64, K = 8 rows; lo*hi, hi*lo, hi*hi) per stage accumulate in TMEM over ALL the
//     stages of the CTA; one read-out and 64 x 64 atomic adds per CTA at the end.  M is 128 because the
//     accumulator layout of M = 128 (lane = row) is the one the rest of the library uses: the upper 64 rows are
//     computed from whatever follows the A tile in shared memory (finite values) and never read.
#include "common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace gtb {

using namespace tc;

constexpr int AT_ROWS = 64;
constexpr int AT_STAGE = 65536;            // A hi 16 KB | A lo 16 KB | B hi 16 KB | B lo 16 KB
constexpr int AT_STAGES = 3;
constexpr int AT_BARS = AT_STAGES * AT_STAGE;  // full[3] | conv[3] | empty[3] | done
constexpr int AT_COLSUM = AT_BARS + 96;    // 64 floats
constexpr int AT_TMEM_SLOT = AT_COLSUM + 256;
constexpr int AT_SMEM = AT_TMEM_SLOT + 16;
constexpr int AT_THREADS = 320;
static_assert(AT_SMEM <= 232448, "shared-memory layout");

__device__ int g_at_fault = 0;

struct AtParams {
  CUtensorMap a_map;  // A [*, 64] fp32: box 32 x 64 (tile mode) or 32 x 1 (gather mode)
  CUtensorMap b_map;  // B [n, 64] fp32: box 32 x 64
  const int32_t* a_index;
  float* out;
  float* colsum;
  int64_t n_rows;
  int32_t n_tiles, out_ld, a_relu;
};

__device__ __noinline__ void at_timeout() {
  atomicExch(&g_at_fault, 1);
  __trap();
}
__device__ __forceinline__ void at_wait(uint32_t bar, uint32_t parity) {
  {  // first try outside the loop: loads issued in front of the wait stay in flight under one blocking try (see edge_ws.cu)
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
#pragma unroll 1
  for (uint32_t i = 0; i < 20000000u; ++i) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  at_timeout();
}

// MN-major tf32 operand.  The only shared-memory layout the tensor core takes for 32-bit MN-major operands is
// SWIZZLE_128B_BASE32B (cutlass sm100_smem_selector: "for mn-major tf32 operands, SW128_32B is the only available
// smem layout"): atoms of 32 elements (128 bytes) along M / N x 4 rows along K, 32-byte granules XOR-ed with the row
// inside the atom -- what a TMA load with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes for a [rows][32 columns] box.
// Atoms along M / N are `lbo` bytes apart (the two K tiles of a 64-column block), 4-row groups along K 512 bytes.
__device__ __forceinline__ uint64_t at_desc(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)1 << 61;   // SWIZZLE_128B_BASE32B
  return d;
}

__global__ void __launch_bounds__(AT_THREADS, 1) rows_atb_tc_kernel(const __grid_constant__ AtParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sm0 = smem_u32(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + AT_TMEM_SLOT);
  if (sm0 & 1023u) {
    if (tid == 0) atomicExch(&g_at_fault, 3);
    __trap();
  }
  const uint32_t full = sm0 + AT_BARS, conv = full + 24, empty = full + 48, done = full + 72;
  if (tid == 0) {
    for (int s = 0; s < AT_STAGES; ++s) {
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + AT_BARS + 8 * s), 1);       // full: the producer's expect_tx
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + AT_BARS + 24 + 8 * s), 8);  // conv: one arrive per converter warp
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + AT_BARS + 48 + 8 * s), 1);  // empty: tcgen05.commit
    }
    mbar_init(reinterpret_cast<uint64_t*>(smem_raw + AT_BARS + 72), 1);
    fence_barrier_init();
  }
  if (tid < 64) reinterpret_cast<float*>(smem_raw + AT_COLSUM)[tid] = 0.f;
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = *tmem_slot;

  // tiles of this CTA: blockIdx.x, + gridDim.x, ...
  const int g = (int)gridDim.x;
  const int n_it = (int)blockIdx.x < p.n_tiles ? (p.n_tiles - (int)blockIdx.x + g - 1) / g : 0;

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      tma::prefetch_map(&p.a_map);
      tma::prefetch_map(&p.b_map);
    }
    for (int it = 0; it < n_it; ++it) {
      const int s = it % AT_STAGES;
      const uint32_t row0 = (uint32_t)((int)blockIdx.x + it * g) * AT_ROWS;
      const int rows_here = (int)min((int64_t)AT_ROWS, p.n_rows - (int64_t)row0);
      const uint32_t st = sm0 + s * AT_STAGE;
      if (it >= AT_STAGES) at_wait(empty + 8 * s, (uint32_t)(it / AT_STAGES - 1) & 1u);
      if (lane == 0) tma::mbar_expect_tx(full + 8 * s, 32768);
      __syncwarp();
      if (p.a_index != nullptr) {
        // rows 2 lane, 2 lane + 1 of the tile; rows past the end gather row 0 (their B rows arrive as zeros)
        int2 r2;
        r2.x = 2 * lane + 0 < rows_here ? __ldg(p.a_index + row0 + 2 * lane + 0) : 0;
        r2.y = 2 * lane + 1 < rows_here ? __ldg(p.a_index + row0 + 2 * lane + 1) : 0;
#pragma unroll 4
        for (int j = 0; j < 16; ++j) {
          const int a = __shfl_sync(0xffffffffu, r2.x, 2 * j), b = __shfl_sync(0xffffffffu, r2.y, 2 * j);
          const int c = __shfl_sync(0xffffffffu, r2.x, 2 * j + 1), d = __shfl_sync(0xffffffffu, r2.y, 2 * j + 1);
          if (elect_one()) {
            tma::gather4(st + j * 512, &p.a_map, full + 8 * s, 0, a, b, c, d);
            tma::gather4(st + 8192 + j * 512, &p.a_map, full + 8 * s, 32, a, b, c, d);
          }
          __syncwarp();
        }
      }
      if (elect_one()) {
        if (p.a_index == nullptr) {
          tma::load_2d(st, &p.a_map, full + 8 * s, 0, (int)row0);  // rows past the end of the table arrive as zeros
          tma::load_2d(st + 8192, &p.a_map, full + 8 * s, 32, (int)row0);
        }
        tma::load_2d(st + 32768, &p.b_map, full + 8 * s, 0, (int)row0);
        tma::load_2d(st + 32768 + 8192, &p.b_map, full + 8 * s, 32, (int)row0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================================================= MMA issue
    // instruction descriptor: tf32 x tf32 -> f32, A and B MN-major (bits 15 / 16), M = 128, N = 64
    const uint32_t idesc = make_idesc_tf32(128, 64) | (1u << 15) | (1u << 16);
    for (int it = 0; it < n_it; ++it) {
      const int s = it % AT_STAGES;
      const uint32_t st = sm0 + s * AT_STAGE;
      at_wait(conv + 8 * s, (uint32_t)(it / AT_STAGES) & 1u);
      tc_fence_after_sync();
      const uint64_t a_hi = at_desc(st, 8192), a_lo = at_desc(st + 16384, 8192);
      const uint64_t b_hi = at_desc(st + 32768, 8192), b_lo = at_desc(st + 49152, 8192);
      if (elect_one()) {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {  // small terms first: lo*hi, hi*lo, hi*hi
          const uint64_t ad = (pass == 0) ? a_lo : a_hi;
          const uint64_t bd = (pass == 1) ? b_lo : b_hi;
#pragma unroll
          for (int kg = 0; kg < 8; ++kg)  // eight rows per step: the next 1024-byte group of both tiles
            mma_tf32_ss(tm, ad + (uint64_t)(kg * 64), bd + (uint64_t)(kg * 64), idesc, it > 0 || pass > 0 || kg > 0);
        }
        mma_commit_addr(empty + 8 * s);
        if (it + 1 == n_it) mma_commit_addr(done);
      }
      __syncwarp();
    }
  } else {
    // ================================================================= converters (256 threads)
    const int t = tid - 64;
    const int c = t & 7;  // logical 16-byte chunk of a 128-byte tile row: columns 4 c .. 4 c + 3 of a K tile
    f32x2 cs[4];          // column sums of B: K tile 0 (two pairs), K tile 1 (two pairs)
#pragma unroll
    for (int i = 0; i < 4; ++i) cs[i] = pack2(0.f, 0.f);
    for (int it = 0; it < n_it; ++it) {
      const int s = it % AT_STAGES;
      const uint32_t st = sm0 + s * AT_STAGE;
      at_wait(full + 8 * s, (uint32_t)(it / AT_STAGES) & 1u);
      // thread t: chunk c of the rows (t >> 3) and (t >> 3) + 32 of both K tiles, of A and of B
#pragma unroll
      for (int m = 0; m < 2; ++m) {      // 0: A, 1: B
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int row = (t >> 3) + 32 * h;
            // logical chunk c of the row: 32-byte granule (c >> 1) ^ (row & 3), 16-byte half c & 1
            const uint32_t a = st + m * 32768 + kt * 8192 + row * 128 + (((((c >> 1) ^ (row & 3)) << 1) | (c & 1)) << 4);
            float4 v = lds128(a);
            if (m == 0 && p.a_relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
            f32x2 h0, l0, h1, l1;
            split_tf32_act2(pack2(v.x, v.y), h0, l0);
            split_tf32_act2(pack2(v.z, v.w), h1, l1);
            if (m == 1) {
              cs[2 * kt] = add2(cs[2 * kt], pack2(v.x, v.y));
              cs[2 * kt + 1] = add2(cs[2 * kt + 1], pack2(v.z, v.w));
            }
            float4 hi, lo;
            unpack2(h0, hi.x, hi.y); unpack2(h1, hi.z, hi.w);
            unpack2(l0, lo.x, lo.y); unpack2(l1, lo.z, lo.w);
            sts128(a, hi);
            sts128(a + 16384, lo);
          }
        }
      }
      fence_proxy_async_smem();  // the tensor core reads the tiles through the async proxy
      __syncwarp();
      if (lane == 0) tma::mbar_arrive(conv + 8 * s);
    }
    if (p.colsum != nullptr) {
      float* sc = reinterpret_cast<float*>(smem_raw + AT_COLSUM);
#pragma unroll
      for (int kt = 0; kt < 2; ++kt) {
        float a, b, cc, d;
        unpack2(cs[2 * kt], a, b);
        unpack2(cs[2 * kt + 1], cc, d);
        atomicAdd(sc + 32 * kt + 4 * c + 0, a);
        atomicAdd(sc + 32 * kt + 4 * c + 1, b);
        atomicAdd(sc + 32 * kt + 4 * c + 2, cc);
        atomicAdd(sc + 32 * kt + 4 * c + 3, d);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");  // converters only
    if (p.colsum != nullptr && t < 64) atomicAdd(p.colsum + t, reinterpret_cast<float*>(smem_raw + AT_COLSUM)[t]);
    if (n_it > 0 && (warp == 4 || warp == 5)) {
      // accumulator rows 0 .. 63 (TMEM lanes 32 (warp % 4) + lane): out[k][n] += D[k][n]
      at_wait(done, 0);
      tc_fence_after_sync();
      const int k = 32 * (warp & 3) + lane;
      float* orow = p.out + (size_t)k * p.out_ld;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t acc[16];
        tmem_ld16(tm + ((uint32_t)(32 * (warp & 3)) << 16) + 16 * q, acc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) atomicAdd(orow + 16 * q + j, __uint_as_float(acc[j]));
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

// Takes the launch when both blocks are 64 wide and TMA-addressable; *handled = false leaves it to the CUDA-core kernel.
int rows_atb_tc(const float* A, int a_ld, const int32_t* a_index, int a_relu, int ka, const float* B, int b_ld, int nb,
                int64_t n_rows, float* out, int out_ld, float* colsum, cudaStream_t st, bool* handled) {
  *handled = false;
  static const bool disabled = getenv("GTB_NO_ATB_TC") != nullptr;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (disabled || ka != 64 || nb != 64 || (a_ld & 3) || (b_ld & 3) || !al16(A) || !al16(B) || n_rows < 4096 ||
      n_rows >= (1ll << 31) - 256 || tma::encode_fn() == nullptr)
    return GTB_OK;
  AtParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t a_rows = a_index != nullptr ? (1ull << 31) : (uint64_t)n_rows;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  if (!tma::make_map_2d(&p.a_map, A, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a_rows, 64, (uint64_t)a_ld, 32, a_index != nullptr ? 1 : AT_ROWS, sw) ||
      !tma::make_map_2d(&p.b_map, B, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)n_rows, 64, (uint64_t)b_ld, 32, AT_ROWS, sw))
    return GTB_OK;
  p.a_index = a_index;
  p.out = out;
  p.colsum = colsum;
  p.n_rows = n_rows;
  p.n_tiles = (int32_t)((n_rows + AT_ROWS - 1) / AT_ROWS);
  p.out_ld = out_ld;
  p.a_relu = a_relu;
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(rows_atb_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(rows_atb_tc)");
    configured = true;
  }
  // every CTA ends with 64 x 64 atomics onto the same addresses: at least 16 stages per CTA
  const int grid = (int)imin64(((int64_t)p.n_tiles + 15) / 16, (int64_t)kNumSMs);
  rows_atb_tc_kernel<<<grid, AT_THREADS, AT_SMEM, st>>>(p);
  GTB_CHECK_LAUNCH("rows_atb_tc_kernel");
  *handled = true;
  return GTB_OK;
}

int at_fault_flag(int* out) { return check_cuda(cudaMemcpyFromSymbol(out, g_at_fault, sizeof(int)), "cudaMemcpyFromSymbol(g_at_fault)"); }

}  // namespace gtb
