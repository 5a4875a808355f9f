// Edge encoder of the edge classifier in one warp-specialised launch (sm_100a):
//
//   out[r] = act(W1 relu(W0 x[i(r)] + b0) + b1)          x: [E, 4] raw edge features, W0: [64, 4], W1: [64, 64]
//
// i.e. reference models/edge_classifier.py:103 (`relu(ec_edge_encoder(edge_attr))`, MLP of models/mlp.py:18-62 with
// L = 2) with the rows gathered through the plan's `perm`, so that the stack behind it streams its edge features
// in destination-sorted order.  16 bytes in, 256 bytes out per edge: a write-bound launch.  The generic tiles
// (mlp_tc.cu) push the K = 4 first Linear through a tf32 split, a TMEM round trip and 24 MMAs like any other;
// here it runs on the CUDA cores (4 FMAs per output, exact fp32), only the 64 x 64 Linear goes to the tensor
// core (3xTF32, A operand in TMEM), and the machine mapping is the edge kernel's (edge_ws.cu):
//
//   * one persistent CTA per SM, two tiles ("contexts") in flight, 20 warps: 16 row-owner warps (thread = row x
//     16-column quarter) that alternate between the contexts, one TMA producer and one MMA-issue warp per context;
//   * producer: the tile's 128 feature rows by tile::gather4 (16-byte box rows, four row coordinates per
//     instruction), two buffers per context;
//   * owners: first Linear + ReLU + tf32 split -> TMEM; later accumulator -> activation -> output tile in shared
//     memory (the 128-byte-swizzle image); then each warp stores its 8-row band of the tile (2-D TMA store);
//   * every hand-over is an mbarrier: full_x / x_free (producer <-> owners, per buffer), a_ready / d_ready
//     (owners <-> MMA warp), out_ready (row owners -> band owners), out_free (band stored -> rows may be rewritten).
#include "common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace gtb {

using namespace tc;

constexpr int EN_TM = 128;
constexpr int EN_W1 = 0;                    // hi 16 KB | lo 16 KB | bias row (gtb_mlp_pack image of one Linear {64, 64})
constexpr int EN_W1_BYTES = 32768 + 256;
constexpr int EN_W0T = 33792;               // W0 transposed [4][64] fp32, then b0 [64]
constexpr int EN_X = 35840;                 // x buffers: context c, buffer b at EN_X + (2 c + b) * 4096: rows 4 j .. 4 j + 3 (16 bytes each)
                                            // in the first 64 bytes of the 128-byte line j (a TMA destination is 128-byte aligned)
constexpr int EN_XBUF = 4096;
constexpr int EN_OUT = 53248;               // output tiles: context c, buffer b at EN_OUT + (2 c + b) * 32768 (1024-byte aligned)
constexpr int EN_BARS = EN_OUT + 4 * 32768; // per context 64 bytes: full_x[2] | x_free[2] | a_ready | d_ready | out_ready | (pad); then out_free[c][b], weights
constexpr int EN_TMEM_SLOT = EN_BARS + 128 + 32 + 8;
constexpr int EN_SMEM = EN_TMEM_SLOT + 8;
constexpr int EN_THREADS = 640;
constexpr uint32_t EN_A_HI = 0, EN_A_LO = 64, EN_D = 128, EN_CTX = 192;
static_assert((EN_OUT & 1023) == 0 && (EN_X & 127) == 0 && EN_W1_BYTES <= EN_W0T && EN_W0T + 1280 <= EN_X && EN_X + 4 * EN_XBUF <= EN_OUT,
              "shared-memory layout");
static_assert(EN_SMEM <= 232448, "shared-memory layout");

__device__ int g_en_fault = 0;

struct EnParams {
  CUtensorMap x_map;     // x [E, 4] fp32: box 4 x 1 (tile::gather4), no swizzle
  CUtensorMap out_map;   // out [E, 64] fp32: box 32 x 8, 128-byte swizzle
  const int32_t* index;  // perm or nullptr
  const float* w0;       // [64, 4]
  const float* b0;       // [64] or nullptr
  const unsigned char* packed_w1;
  int64_t n_rows;
  int32_t n_tiles, final_relu;
};

__device__ __noinline__ void en_timeout() {
  atomicExch(&g_en_fault, 1);
  __trap();
}
__device__ __forceinline__ void en_wait(uint32_t bar, uint32_t parity) {
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
  en_timeout();
}

struct EnOwner {
  const EnParams& p;
  uint32_t sm0;
  int w, lane, r, qd;
  uint32_t rx, own;  // swizzle term of the own row, own row inside the own K tile of an output tile
  int tile00, nA, nB;

  __device__ __forceinline__ int n_of(int c) const { return c ? nB : nA; }
  __device__ __forceinline__ int tile_of(int c, int t) const { return tile00 + (2 * t + c) * (int)gridDim.x; }
  __device__ __forceinline__ uint32_t bar(int c, int which) const { return sm0 + EN_BARS + 64 * c + 8 * which; }
  __device__ __forceinline__ uint32_t out_free(int c, int b) const { return sm0 + EN_BARS + 128 + 8 * (2 * c + b); }
  __device__ __forceinline__ uint32_t tm_lane(int c) const { return (uint32_t)c * EN_CTX + ((uint32_t)((w & 3) * 32) << 16); }
  __device__ __forceinline__ uint32_t chunk(int q) const { return (((uint32_t)(4 * (qd & 1) + q)) << 4) ^ rx; }
  __device__ __forceinline__ void arrive(uint32_t b) {
    __syncwarp();
    if (lane == 0) tma::mbar_arrive(b);
  }

  // ---- C0: own row's four features -> first Linear on the CUDA cores (own 16 outputs) -> ReLU -> tf32 hi / lo -> TMEM
  __device__ __forceinline__ void c0(int c, int t) {
    const int b = t & 1;
    en_wait(bar(c, b), (uint32_t)(t >> 1) & 1u);
    const float4 x = lds128(sm0 + EN_X + (2 * c + b) * EN_XBUF + 128 * (r >> 2) + 16 * (r & 3));
    const uint32_t wa = sm0 + EN_W0T + 64 * qd;
    f32x2 acc[8];
    {
      const uint32_t ba = sm0 + EN_W0T + 1024 + 64 * qd;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 bb = lds128(ba + 16 * q);
        acc[2 * q] = pack2(bb.x, bb.y);
        acc[2 * q + 1] = pack2(bb.z, bb.w);
      }
    }
    const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const f32x2 xk = pack2(xs[k], xs[k]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 wv = lds128(wa + 256 * k + 16 * q);  // the same address across the warp: a broadcast
        acc[2 * q] = fma2(pack2(wv.x, wv.y), xk, acc[2 * q]);
        acc[2 * q + 1] = fma2(pack2(wv.z, wv.w), xk, acc[2 * q + 1]);
      }
    }
    float v[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      unpack2(acc[j], v[2 * j], v[2 * j + 1]);
      v[2 * j] = fmaxf(v[2 * j], 0.f);
      v[2 * j + 1] = fmaxf(v[2 * j + 1], 0.f);
    }
    split_store16(tm_lane(c) + EN_A_HI + 16 * qd, tm_lane(c) + EN_A_LO + 16 * qd, v);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) {
      tma::mbar_arrive(bar(c, 4));      // a_ready
      tma::mbar_arrive(bar(c, 2 + b));  // x_free: the producer may refill this buffer
    }
  }

  // ---- E2: accumulator + b1 -> activation -> output tile in shared memory
  __device__ __forceinline__ void e2(int c, int t) {
    const int b = t & 1;
    const uint32_t sl = sm0 + EN_OUT + (2 * c + b) * 32768;
    if (t >= 2) en_wait(out_free(c, b), (uint32_t)((t >> 1) - 1) & 1u);  // every band of this buffer has been stored
    en_wait(bar(c, 5), (uint32_t)t & 1u);
    tc_fence_after_sync();
    uint32_t acc[16];
    tmem_ld16(tm_lane(c) + EN_D + 16 * qd, acc);
    tmem_ld_wait();
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
    const uint32_t ba = sm0 + EN_W1 + 32768 + 64 * qd;
    add16(v, lds128(ba), lds128(ba + 16), lds128(ba + 32), lds128(ba + 48));
    if (p.final_relu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) sts128(sl + own + chunk(q), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    tc_fence_before_sync();
    fence_proxy_async_smem();  // the TMA store reads the tile through the async proxy
    arrive(bar(c, 6));
  }

  // ---- ST: warp w stores rows 8 w .. 8 w + 7 of the tile (all 64 columns)
  __device__ __forceinline__ void st(int c, int t) {
    const int b = t & 1;
    const uint32_t sl = sm0 + EN_OUT + (2 * c + b) * 32768;
    const uint32_t row0 = (uint32_t)tile_of(c, t) * EN_TM;
    const int rows_here = (int)min((int64_t)EN_TM, p.n_rows - (int64_t)row0);
    en_wait(bar(c, 6), (uint32_t)t & 1u);
    if (8 * w < rows_here) {
      if (elect_one()) {  // rows past the end of the table are clipped
        tma::store_2d(&p.out_map, sl + (uint32_t)(8 * w) * 128u, 0, (int)row0 + 8 * w);
        tma::store_2d(&p.out_map, sl + 16384u + (uint32_t)(8 * w) * 128u, 32, (int)row0 + 8 * w);
      }
      __syncwarp();
    }
    tma::bulk_commit();
  }
  // the stores of tile iteration t have read their bands: buffer t & 1 may be rewritten
  __device__ __forceinline__ void released(int c, int t) {
    tma::bulk_wait_read0();
    arrive(out_free(c, t & 1));
  }
};

__global__ void __launch_bounds__(EN_THREADS, 1) edge_encoder_ws_kernel(const __grid_constant__ EnParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sm0 = smem_u32(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + EN_TMEM_SLOT);
  if (sm0 & 1023u) {
    if (tid == 0) atomicExch(&g_en_fault, 3);
    __trap();
  }
  const uint32_t wbar = sm0 + EN_BARS + 128 + 32;
  if (tid == 0) {
    for (int c = 0; c < 2; ++c) {
      unsigned char* b = smem_raw + EN_BARS + 64 * c;
      mbar_init(reinterpret_cast<uint64_t*>(b + 0), 1);    // full_x[0]: the producer's expect_tx
      mbar_init(reinterpret_cast<uint64_t*>(b + 8), 1);    // full_x[1]
      mbar_init(reinterpret_cast<uint64_t*>(b + 16), 16);  // x_free[0]: one arrive per owner warp
      mbar_init(reinterpret_cast<uint64_t*>(b + 24), 16);  // x_free[1]
      mbar_init(reinterpret_cast<uint64_t*>(b + 32), 16);  // a_ready
      mbar_init(reinterpret_cast<uint64_t*>(b + 40), 1);   // d_ready
      mbar_init(reinterpret_cast<uint64_t*>(b + 48), 16);  // out_ready
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + EN_BARS + 128 + 8 * (2 * c)), 16);      // out_free[c][0]
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + EN_BARS + 128 + 8 * (2 * c + 1)), 16);  // out_free[c][1]
    }
    mbar_init(reinterpret_cast<uint64_t*>(smem_raw + EN_BARS + 128 + 32), 1);
    fence_barrier_init();
    tma::mbar_expect_tx(wbar, EN_W1_BYTES);
    tma::bulk_g2s(sm0 + EN_W1, p.packed_w1, EN_W1_BYTES, wbar);
  }
  if (tid < 256) {  // W0 [64][4] -> transposed [4][64]; b0
    const int j = tid >> 2, k = tid & 3;
    reinterpret_cast<float*>(smem_raw + EN_W0T)[k * 64 + j] = __ldg(p.w0 + tid);
    if (tid < 64) reinterpret_cast<float*>(smem_raw + EN_W0T + 1024)[tid] = p.b0 ? __ldg(p.b0 + tid) : 0.f;
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (*tmem_slot != 0) {
    if (tid == 0) atomicExch(&g_en_fault, 2);
    __trap();
  }

  const int g = (int)gridDim.x;
  int tile0[2], n_t[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    tile0[c] = (int)blockIdx.x + g * c;
    n_t[c] = tile0[c] < p.n_tiles ? (p.n_tiles - tile0[c] + 2 * g - 1) / (2 * g) : 0;
  }

  if (warp < 16) {
    // ================================================================= row owners
    EnOwner o{p, sm0};
    o.w = warp; o.lane = lane; o.r = 32 * (warp & 3) + lane; o.qd = warp >> 2;
    o.rx = (uint32_t)(o.r & 7) << 4;
    o.own = (uint32_t)(o.qd >> 1) * 16384u + (uint32_t)o.r * 128u;
    o.tile00 = tile0[0]; o.nA = n_t[0]; o.nB = n_t[1];
    en_wait(wbar, 0);  // W1 and its bias row have landed
    // per round: C0 C0' | E2 ST C0+ | E2' ST' C0'+  -- the Linear of one context runs under the stages of the other
#pragma unroll 1
    for (int c = 0; c < 2; ++c)
      if (o.n_of(c) > 0) o.c0(c, 0);
#pragma unroll 1
    for (int t = 0; t < o.nA; ++t) {
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        if (t < o.n_of(c)) {
          o.e2(c, t);
          if (t >= 1) o.released(c, t - 1);
          o.st(c, t);
        }
        if (t + 1 < o.n_of(c)) o.c0(c, t + 1);
      }
    }
    tma::bulk_wait_all0();
  } else if (warp < 18) {
    // ================================================================= TMA producer of context `ctx`
    const int ctx = warp - 16;
    const uint32_t bars = sm0 + EN_BARS + 64 * ctx;
    if (lane == 0) {
      tma::prefetch_map(&p.x_map);
      tma::prefetch_map(&p.out_map);
    }
    for (int t = 0; t < n_t[ctx]; ++t) {
      const int b = t & 1;
      const int tile = tile0[ctx] + t * 2 * g;
      const uint32_t row0 = (uint32_t)tile * EN_TM;
      const int rows_here = (int)min((int64_t)EN_TM, p.n_rows - (int64_t)row0);
      const uint32_t dst = sm0 + EN_X + (2 * ctx + b) * EN_XBUF;
      const uint32_t full_x = bars + 8 * b, x_free = bars + 16 + 8 * b;
      if (t >= 2) en_wait(x_free, (uint32_t)((t >> 1) - 1) & 1u);
      if (lane == 0) tma::mbar_expect_tx(full_x, 2048);
      __syncwarp();
      int4 r4;  // rows past the end gather row 0 (never stored); without an index the rows are the tile's own
      if (p.index != nullptr && rows_here == EN_TM) {
        r4 = __ldg(reinterpret_cast<const int4*>(p.index + row0) + lane);
      } else {
        int v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = 4 * lane + i;
          v[i] = rr < rows_here ? (p.index != nullptr ? __ldg(p.index + row0 + rr) : (int)row0 + rr) : 0;
        }
        r4 = make_int4(v[0], v[1], v[2], v[3]);
      }
#pragma unroll 4
      for (int j = 0; j < 32; ++j) {
        const int a = __shfl_sync(0xffffffffu, r4.x, j), bb = __shfl_sync(0xffffffffu, r4.y, j);
        const int cc = __shfl_sync(0xffffffffu, r4.z, j), d = __shfl_sync(0xffffffffu, r4.w, j);
        if (elect_one()) tma::gather4(dst + j * 128, &p.x_map, full_x, 0, a, bb, cc, d);
        __syncwarp();
      }
    }
  } else {
    // ================================================================= MMA issue of context `ctx`
    const int ctx = warp - 18;
    const uint32_t a_ready = sm0 + EN_BARS + 64 * ctx + 32, d_ready = a_ready + 8;
    const uint32_t tmc = (uint32_t)ctx * EN_CTX;
    const uint32_t idesc = make_idesc_tf32(EN_TM, 64);
    en_wait(wbar, 0);
    const uint64_t bd_hi = make_smem_desc_sw128(sm0 + EN_W1), bd_lo = make_smem_desc_sw128(sm0 + EN_W1 + 16384u);
    for (int t = 0; t < n_t[ctx]; ++t) {
      en_wait(a_ready, (uint32_t)t & 1u);
      tc_fence_after_sync();
      if (elect_one()) {
        bool acc = false;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {  // small terms first: lo*hi, hi*lo, hi*hi
          const uint32_t a = tmc + ((pass == 0) ? EN_A_LO : EN_A_HI);
          const uint64_t bd = (pass == 1) ? bd_lo : bd_hi;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            mma_tf32_ts(tmc + EN_D, a + 8 * ks, bd + (uint64_t)((ks >> 2) * 512 + (ks & 3) * 2), idesc, acc);
            acc = true;
          }
        }
        mma_commit_addr(d_ready);
      }
      __syncwarp();
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0u, 512);
}

int edge_encoder_ws(const float* x, int32_t x_ld, const int32_t* index, int64_t n_rows, int64_t x_rows, const float* w0,
                    const float* b0, const void* packed_w1, int32_t final_relu, float* out, int32_t out_ld, cudaStream_t st) {
  GTB_REQUIRE(x && w0 && packed_w1 && out, GTB_ERR_BAD_ARG, "gtb_edge_encoder_f32: null argument");
  GTB_REQUIRE(n_rows >= 0 && n_rows < (1ll << 31) - 256 && x_rows >= 0, GTB_ERR_BAD_ARG, "gtb_edge_encoder_f32: bad row count");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  GTB_REQUIRE(al16(x) && al16(out) && al16(packed_w1) && al16(w0) && (index == nullptr || al16(index)), GTB_ERR_BAD_ARG,
              "gtb_edge_encoder_f32: pointers must be 16-byte aligned");
  GTB_REQUIRE(x_ld >= 4 && !(x_ld & 3) && out_ld >= 64 && !(out_ld & 3), GTB_ERR_BAD_ARG,
              "gtb_edge_encoder_f32: row strides must be multiples of 4 elements (x: 4 columns, out: 64)");
  GTB_REQUIRE(tma::encode_fn() != nullptr, GTB_ERR_CUDA, "gtb_edge_encoder_f32: cuTensorMapEncodeTiled is not available");
  if (n_rows == 0) return GTB_OK;
  EnParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t xr = index != nullptr ? (uint64_t)(x_rows > 0 ? x_rows : (1ll << 31)) : (uint64_t)n_rows;
  GTB_REQUIRE(tma::make_map_2d(&p.x_map, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, xr, 4, (uint64_t)x_ld, 4, 1, CU_TENSOR_MAP_SWIZZLE_NONE) &&
                  tma::make_map_2d(&p.out_map, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)n_rows, 64, (uint64_t)out_ld, 32, 8),
              GTB_ERR_CUDA, "gtb_edge_encoder_f32: the driver refused a tensor map");
  p.index = index;
  p.w0 = w0;
  p.b0 = b0;
  p.packed_w1 = static_cast<const unsigned char*>(packed_w1);
  p.n_rows = n_rows;
  p.n_tiles = (int32_t)((n_rows + EN_TM - 1) / EN_TM);
  p.final_relu = final_relu;
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(edge_encoder_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EN_SMEM);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(edge_encoder_ws)");
    configured = true;
  }
  const int pairs = (p.n_tiles + 1) / 2;
  const int grid = pairs < kNumSMs ? pairs : kNumSMs;
  edge_encoder_ws_kernel<<<grid, EN_THREADS, EN_SMEM, st>>>(p);
  GTB_CHECK_LAUNCH("edge_encoder_ws_kernel");
  return GTB_OK;
}

int en_fault_flag(int* out) { return check_cuda(cudaMemcpyFromSymbol(out, g_en_fault, sizeof(int)), "cudaMemcpyFromSymbol(g_en_fault)"); }

}  // namespace gtb
