// Node side of one Interaction-Network layer in ONE launch (sm_100a), the 64-wide "wide" shape:
//
//   x_out = res_a * res + res_b * MLP_obj(cat[act(x), aggr])        (reference models/interaction_network.py:92-103,
//                                                                    sqconvex_combination of models/resin.py:17-42)
//   P_a   = act'(x_out) Wa^T ,  P_b = act'(x_out) Wb^T              (the per-node products the NEXT consumer's first
//                                                                    Linear gathers: the next layer's relational model,
//                                                                    interaction_network.py:75-89, or the W head,
//                                                                    edge_classifier.py:108-117; GTB_SRC_PROJECTED)
//   aggr  = 0                                                        (optional: the aggregate is handed back zeroed to
//                                                                    the next layer's edge kernel)
//
// Without this kernel a layer's node side is four launches over the same 100k rows (object model, two
// pre-projections, zero-fill), each paying its own prologue for 5 tiles per SM.  Same arithmetic as the other
// tensor-core tiles (3xTF32, A operand in TMEM, fp32 accumulation), mapped like the edge kernel of edge_ws.cu:
//
//   * one persistent CTA per SM, two tiles ("contexts") in flight, 18 warps: warps 0-15 are 512 row owners
//     (thread = row x 16-column quarter) that ping-pong between the contexts stage by stage, warps 16 / 17 issue
//     the tcgen05.mma groups of context 0 / 1;
//   * five MMA groups per tile -- first Linear over x (K = 64), over aggr (K = 64, accumulating), second and third
//     Linear, the two projections (2 x N = 64 into a 128-column accumulator) -- each handed over by two
//     mbarriers (a_ready: owners -> MMA warp, d_ready: tcgen05.commit -> owners);
//   * all weights stay resident: object model 128 KB (the gtb_mlp_pack image for blocks {64, 64}) + two
//     projections 2 x 32 KB.  The rows are N-sized tables that sit in L2 (the edge kernel has just written the
//     aggregate): owners read and write their 64-byte row pieces directly, no staging slots.
//   * projection-only mode (packed_obj == NULL): P_a / P_b of act'(x), for the first layer of a stack.
#include "common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace gtb {

using namespace tc;

constexpr int NW_TM = 128;
constexpr int NW_OBJ = 0;                   // object-model image: L0 hi 32 KB | L0 lo 32 KB | L1 hi | L1 lo | L2 hi | L2 lo | 3 bias rows
constexpr int NW_OBJ_BYTES = 131072 + 768;
constexpr int NW_OBJ_BIAS = 131072;
constexpr int NW_PA = 132096;               // projection image: hi 16 KB | lo 16 KB | (zero) bias row
constexpr int NW_PB = NW_PA + 33792;
constexpr int NW_BARS = NW_PB + 32768;      // per context 16 bytes: a_ready | d_ready; then the weight-arrival barrier
constexpr int NW_TMEM_SLOT = NW_BARS + 40;
constexpr int NW_STAGE = NW_PB + 33792;     // 16 row-owner warps x 2 KB
constexpr int NW_SMEM = NW_STAGE + 16 * 2048;
static_assert(NW_TMEM_SLOT + 4 <= NW_STAGE, "shared-memory layout");
constexpr int NW_THREADS = 576;
constexpr uint32_t NW_A_HI = 0, NW_A_LO = 64, NW_D = 128, NW_CTX = 256;
static_assert(NW_OBJ_BYTES <= NW_PA && (NW_PA & 1023) == 0 && (NW_PB & 1023) == 0, "weight tiles are 1024-byte aligned");
static_assert(NW_SMEM <= 232448, "shared-memory layout");

__device__ int g_nw_fault = 0;

struct NwParams {
  const float* x;
  float* aggr;
  const float* res;
  float* x_out;
  float* pa;
  float* pb;
  const unsigned char* packed_obj;
  const unsigned char* packed_pa;
  const unsigned char* packed_pb;
  int64_t n_rows;
  int32_t n_tiles, x_ld, aggr_ld, res_ld, xo_ld, pa_ld, pb_ld;
  int32_t relu_x, zero_aggr, proj_relu;
  float res_a, res_b;
};

__device__ __noinline__ void nw_timeout() {
  atomicExch(&g_nw_fault, 1);
  __trap();
}
__device__ __forceinline__ void nw_wait(uint32_t bar, uint32_t parity) {
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
  nw_timeout();
}
__device__ __forceinline__ void nw_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <bool OBJ, bool PROJ>
struct NwOwner {
  static constexpr int G = (OBJ ? 4 : 0) + (PROJ ? 1 : 0);  // MMA groups (= accumulator completions) per tile
  const NwParams& p;
  uint32_t sm0;
  int w, lane, r, qd;
  int tile00, nA, nB;
  uint32_t stg;  // this warp's 2 KB staging block: [32 rows][64 bytes], 16-byte chunk k of row i at chunk k ^ (i >> 1)

  __device__ __forceinline__ int n_of(int c) const { return c ? nB : nA; }
  __device__ __forceinline__ int tile_of(int c, int t) const { return tile00 + (2 * t + c) * (int)gridDim.x; }
  __device__ __forceinline__ uint32_t bar(int c, int which) const { return sm0 + NW_BARS + 16 * c + 8 * which; }
  __device__ __forceinline__ uint32_t tm_lane(int c) const { return (uint32_t)c * NW_CTX + ((uint32_t)((w & 3) * 32) << 16); }
  __device__ __forceinline__ uint32_t sw(int row, int chunk) const { return stg + (uint32_t)row * 64u + (uint32_t)((chunk ^ (row >> 1)) & 3) * 16u; }
  __device__ __forceinline__ void a_done(int c) {
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) nw_arrive(bar(c, 0));
  }
  // completion number g of tile iteration t (G per tile)
  __device__ __forceinline__ void wait_d(int c, int t, int g) {
    nw_wait(bar(c, 1), (uint32_t)(t * G + g) & 1u);
    tc_fence_after_sync();
  }
  __device__ __forceinline__ void acc_load(int c, uint32_t col, float (&v)[16]) {
    uint32_t acc[16];
    tmem_ld16(tm_lane(c) + NW_D + col, acc);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
  }
  __device__ __forceinline__ void bias_add(float (&v)[16], int layer) {
    const uint32_t ba = sm0 + NW_OBJ + NW_OBJ_BIAS + 256 * layer + 64 * qd;
    add16(v, lds128(ba), lds128(ba + 16), lds128(ba + 32), lds128(ba + 48));
  }
  // The warp owns the block rows 32 (w & 3) .. + 31 x columns 16 qd .. + 15 of a tile (thread = one row of it, as its
  // TMEM lane dictates).  Global memory is touched in the coalesced shape instead -- lane l moves chunk l & 3 of the
  // rows (l >> 2) + 8 i: eight full 64-byte row pieces per instruction -- and the block turns through the staging block.
  __device__ __forceinline__ float* blk_ptr(float* base, int32_t ld, int tile, int i, bool& ok) const {
    const int64_t row = (int64_t)tile * NW_TM + 32 * (w & 3) + 8 * i + (lane >> 2);
    ok = row < p.n_rows;
    return base + row * (int64_t)ld + 16 * qd + 4 * (lane & 3);
  }
  // Both contexts' blocks in ONE straight-line batch of eight loads: no branch between them (ptxas parks the
  // scoreboard wait of a load at the next branch, so a guarded load per row or per context costs one full memory
  // latency EACH -- measured: 173 us per launch against ~40).  Rows past the end of the table and the tile of a
  // context that has run out are clamped onto valid rows instead; their results are never stored.
  template <bool NC>
  __device__ __forceinline__ void blk_load2(const float* base, int32_t ld, int t, float4 (&g0)[4], float4 (&g1)[4]) const {
    const int tl0 = tile_of(0, t), tl1 = t < nB ? tile_of(1, t) : tl0;
    const float4* q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int64_t row = (int64_t)((i < 4) ? tl0 : tl1) * NW_TM + 32 * (w & 3) + 8 * (i & 3) + (lane >> 2);
      row = row < p.n_rows ? row : p.n_rows - 1;
      q[i] = reinterpret_cast<const float4*>(base + row * (int64_t)ld + 16 * qd + 4 * (lane & 3));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (NC) {
        g0[i] = __ldg(q[i]);
        g1[i] = __ldg(q[4 + i]);
      } else {  // plain (coherent) loads where the same thread writes the address afterwards
        g0[i] = *q[i];
        g1[i] = *q[4 + i];
      }
    }
  }
  __device__ __forceinline__ void blk_zero(float* base, int32_t ld, int tile) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bool ok;
      float4* q = reinterpret_cast<float4*>(blk_ptr(base, ld, tile, i, ok));
      if (ok) *q = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ __forceinline__ void blk_to_row(const float4 (&g)[4], float (&v)[16]) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) sts128(sw(8 * i + (lane >> 2), lane & 3), g[i]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 a = lds128(sw(lane, k));
      v[4 * k] = a.x; v[4 * k + 1] = a.y; v[4 * k + 2] = a.z; v[4 * k + 3] = a.w;
    }
    __syncwarp();
  }
  __device__ __forceinline__ void row_store(const float (&v)[16], float* base, int32_t ld, int tile) const {
#pragma unroll
    for (int k = 0; k < 4; ++k) sts128(sw(lane, k), make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bool ok;
      float4* q = reinterpret_cast<float4*>(blk_ptr(base, ld, tile, i, ok));
      const float4 a = lds128(sw(8 * i + (lane >> 2), lane & 3));
      if (ok) *q = a;
    }
    __syncwarp();
  }
  __device__ __forceinline__ void to_a(int c, const float (&v)[16]) {
    split_store16(tm_lane(c) + NW_A_HI + 16 * qd, tm_lane(c) + NW_A_LO + 16 * qd, v);
    a_done(c);
  }

  // ---- X: act(x) -> A operand (first Linear over the node features; projection-only: the projected operand).
  // The loads of both contexts are in flight together: one exposed memory latency per stage, not two.
  __device__ __forceinline__ void sx(int t) {
    float4 g0[4], g1[4];
    const bool h1 = t < nB;
    blk_load2<true>(p.x, p.x_ld, t, g0, g1);
    const bool relu = OBJ ? p.relu_x : p.proj_relu;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (c == 0 || h1) {
        float v[16];
        blk_to_row(c ? g1 : g0, v);
        if (relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        to_a(c, v);
      }
    }
  }
  // ---- AG: aggr -> A operand once the first group has read the A columns; the aggregate is handed back zeroed
  __device__ __forceinline__ void sa(int t) {
    float4 g0[4], g1[4];
    const bool h1 = t < nB;
    blk_load2<false>(p.aggr, p.aggr_ld, t, g0, g1);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (c == 0 || h1) {
        float v[16];
        blk_to_row(c ? g1 : g0, v);
        if (p.zero_aggr) blk_zero(p.aggr, p.aggr_ld, tile_of(c, t));  // behind the loads of BOTH contexts
        uint32_t hi[16], lo[16];
        split16(v, hi, lo);  // in front of the wait: only the two TMEM stores are left behind it
        wait_d(c, t, 0);
        tmem_st16(tm_lane(c) + NW_A_HI + 16 * qd, hi);
        tmem_st16(tm_lane(c) + NW_A_LO + 16 * qd, lo);
        a_done(c);
      }
    }
  }
  // ---- hidden layers: D + b -> ReLU -> A operand
  __device__ __forceinline__ void sh(int c, int t, int layer) {
    wait_d(c, t, 1 + layer);
    float v[16];
    acc_load(c, 16 * qd, v);
    bias_add(v, layer);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
    to_a(c, v);
  }
  // ---- output of the object model: residual, store, and the operand of the projections
  __device__ __forceinline__ void so(int t) {
    float4 g0[4], g1[4];
    const bool h1 = t < nB;
    const bool has_res = p.res != nullptr;
    if (has_res) blk_load2<true>(p.res, p.res_ld, t, g0, g1);  // issued in front of the accumulator waits
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (c == 0 || h1) {
        float q[16];
        if (has_res) blk_to_row(c ? g1 : g0, q);
        wait_d(c, t, 3);
        float v[16];
        acc_load(c, 16 * qd, v);
        bias_add(v, 2);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= p.res_b;
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaf(p.res_a, q[j], v[j]);
        }
        row_store(v, p.x_out, p.xo_ld, tile_of(c, t));
        if (PROJ) {
          if (p.proj_relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          to_a(c, v);
        }
      }
    }
  }
  // ---- projections out
  __device__ __forceinline__ void sp(int c, int t) {
    wait_d(c, t, G - 1);
    float a[16];
    acc_load(c, 16 * qd, a);
    row_store(a, p.pa, p.pa_ld, tile_of(c, t));
    acc_load(c, 64 + 16 * qd, a);
    row_store(a, p.pb, p.pb_ld, tile_of(c, t));
    tc_fence_before_sync();  // the next tile's first group overwrites these columns: ordered by the a_ready arrive
  }
};

// one N = 64, K = 64 Linear in 3xTF32 from the A operand in TMEM: 24 MMAs (small terms first: lo*hi, hi*lo, hi*hi)
__device__ __forceinline__ void nw_linear(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi, uint32_t w_lo, uint32_t idesc, bool acc) {
  const uint64_t bd_hi = make_smem_desc_sw128(w_hi), bd_lo = make_smem_desc_sw128(w_lo);
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t a = (pass == 0) ? a_lo : a_hi;
    const uint64_t bd = (pass == 1) ? bd_lo : bd_hi;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      mma_tf32_ts(d, a + 8 * ks, bd + (uint64_t)((ks >> 2) * 512 + (ks & 3) * 2), idesc, acc);
      acc = true;
    }
  }
}

template <bool OBJ, bool PROJ>
__global__ void __launch_bounds__(NW_THREADS, 1) in_node_ws_kernel(const __grid_constant__ NwParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sm0 = smem_u32(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + NW_TMEM_SLOT);
  if (sm0 & 1023u) {
    if (tid == 0) atomicExch(&g_nw_fault, 3);
    __trap();
  }
  const uint32_t wbar = sm0 + NW_BARS + 32;
  if (tid == 0) {
    for (int c = 0; c < 2; ++c) {
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + NW_BARS + 16 * c), 16);     // a_ready: one arrive per owner warp
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + NW_BARS + 16 * c + 8), 1);  // d_ready: tcgen05.commit
    }
    mbar_init(reinterpret_cast<uint64_t*>(smem_raw + NW_BARS + 32), 1);
    fence_barrier_init();
    // resident weights: a handful of bulk copies (async proxy, the proxy the tensor core reads through)
    tma::mbar_expect_tx(wbar, (OBJ ? NW_OBJ_BYTES : 0) + (PROJ ? 2 * 32768 : 0));
    if (OBJ) {
#pragma unroll 1
      for (int i = 0; i < 4; ++i) tma::bulk_g2s(sm0 + NW_OBJ + i * (NW_OBJ_BYTES / 4), p.packed_obj + i * (NW_OBJ_BYTES / 4), NW_OBJ_BYTES / 4, wbar);
    }
    if (PROJ) {
      tma::bulk_g2s(sm0 + NW_PA, p.packed_pa, 32768, wbar);
      tma::bulk_g2s(sm0 + NW_PB, p.packed_pb, 32768, wbar);
    }
  }
  if (warp < 16) {
    // every line of this CTA's x / aggr tiles towards L2 now, under the weight copies: the stage loads below then
    // pay an L2 hit instead of a DRAM round trip (the stages are latency-bound: ~3 tile pairs per CTA).
    // threads 0-255: the 256 lines of the x tile, threads 256-511: the aggr tile
    const int line = tid & 255;
    const float* base = (tid < 256 || !OBJ) ? p.x : p.aggr;
    const int32_t ld = (tid < 256 || !OBJ) ? p.x_ld : p.aggr_ld;
    if (OBJ || tid < 256) {
      for (int tile = (int)blockIdx.x; tile < p.n_tiles; tile += (int)gridDim.x) {
        int64_t row = (int64_t)tile * NW_TM + (line >> 1);
        row = row < p.n_rows ? row : p.n_rows - 1;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + row * (int64_t)ld + 32 * (line & 1)));
      }
    }
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (*tmem_slot != 0) {
    if (tid == 0) atomicExch(&g_nw_fault, 2);
    __trap();
  }
  nw_wait(wbar, 0);  // weights and biases have landed

  const int g = (int)gridDim.x;
  int tile0[2], n_t[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    tile0[c] = (int)blockIdx.x + g * c;
    n_t[c] = tile0[c] < p.n_tiles ? (p.n_tiles - tile0[c] + 2 * g - 1) / (2 * g) : 0;
  }

  if (warp < 16) {
    NwOwner<OBJ, PROJ> o{p, sm0};
    o.w = warp; o.lane = lane; o.r = 32 * (warp & 3) + lane; o.qd = warp >> 2;
    o.tile00 = tile0[0]; o.nA = n_t[0]; o.nB = n_t[1];
    o.stg = sm0 + NW_STAGE + 2048u * (uint32_t)warp;
#pragma unroll 1
    for (int t = 0; t < o.nA; ++t) {
      o.sx(t);
      if (OBJ) {
        o.sa(t);
#pragma unroll 1
        for (int l = 0; l < 2; ++l) {
#pragma unroll 1
          for (int c = 0; c < 2; ++c)
            if (t < o.n_of(c)) o.sh(c, t, l);
        }
        o.so(t);
      }
      if (PROJ) {
#pragma unroll 1
        for (int c = 0; c < 2; ++c)
          if (t < o.n_of(c)) o.sp(c, t);
      }
    }
  } else {
    // ================================================================= MMA issue of context `ctx`
    const int ctx = warp - 16;
    const uint32_t a_ready = sm0 + NW_BARS + 16 * ctx, d_ready = a_ready + 8;
    const uint32_t tmc = (uint32_t)ctx * NW_CTX;
    const uint32_t idesc = make_idesc_tf32(NW_TM, 64);
    constexpr int G = (OBJ ? 4 : 0) + (PROJ ? 1 : 0);
    int n = 0;
    for (int t = 0; t < n_t[ctx]; ++t) {
#pragma unroll 1
      for (int gi = 0; gi < G; ++gi, ++n) {
        nw_wait(a_ready, (uint32_t)n & 1u);
        tc_fence_after_sync();
        if (elect_one()) {
          if (OBJ && gi < 4) {
            // group 0 / 1: the x / aggr half of the first Linear (K tiles 0-1 / 2-3 of its image), 2 / 3: second / third Linear
            const uint32_t hi = sm0 + NW_OBJ + (gi == 0 ? 0u : gi == 1 ? 16384u : gi == 2 ? 65536u : 98304u);
            const uint32_t lo = hi + (gi < 2 ? 32768u : 16384u);
            nw_linear(tmc + NW_D, tmc + NW_A_HI, tmc + NW_A_LO, hi, lo, idesc, gi == 1);
          } else {
            nw_linear(tmc + NW_D, tmc + NW_A_HI, tmc + NW_A_LO, sm0 + NW_PA, sm0 + NW_PA + 16384u, idesc, false);
            nw_linear(tmc + NW_D + 64, tmc + NW_A_HI, tmc + NW_A_LO, sm0 + NW_PB, sm0 + NW_PB + 16384u, idesc, false);
          }
          mma_commit_addr(d_ready);
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0u, 512);
}

template <bool OBJ, bool PROJ>
static cudaError_t nw_configure() {
  return cudaFuncSetAttribute(in_node_ws_kernel<OBJ, PROJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, NW_SMEM);
}

int in_node_ws(const float* x, int32_t x_ld, int32_t relu_x, float* aggr, int32_t aggr_ld, int32_t zero_aggr, int64_t n_nodes,
               const void* packed_obj, float res_a, float res_b, const float* res, int32_t res_ld, float* x_out, int32_t xo_ld,
               const void* packed_pa, const void* packed_pb, int32_t proj_relu, float* p_a, int32_t pa_ld, float* p_b,
               int32_t pb_ld, cudaStream_t st) {
  const bool obj = packed_obj != nullptr, proj = packed_pa != nullptr;
  GTB_REQUIRE(obj || proj, GTB_ERR_BAD_ARG, "gtb_in_node_fused_f32: neither an object model nor projections given");
  GTB_REQUIRE(x != nullptr && n_nodes >= 0 && n_nodes < (1ll << 31) - 256, GTB_ERR_BAD_ARG, "gtb_in_node_fused_f32: bad x / n_nodes");
  GTB_REQUIRE(!obj || (aggr != nullptr && x_out != nullptr), GTB_ERR_BAD_ARG, "gtb_in_node_fused_f32: aggr / x_out missing");
  GTB_REQUIRE(!proj || (packed_pb != nullptr && p_a != nullptr && p_b != nullptr), GTB_ERR_BAD_ARG,
              "gtb_in_node_fused_f32: projections come in pairs (packed_pb, p_a, p_b)");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  GTB_REQUIRE(al16(x) && al16(aggr) && al16(res) && al16(x_out) && al16(p_a) && al16(p_b) && al16(packed_obj) && al16(packed_pa) &&
                  al16(packed_pb),
              GTB_ERR_BAD_ARG, "gtb_in_node_fused_f32: pointers must be 16-byte aligned");
  auto ld_ok = [](int32_t ld) { return ld >= 64 && (ld & 3) == 0; };
  GTB_REQUIRE(ld_ok(x_ld) && (!obj || (ld_ok(aggr_ld) && ld_ok(xo_ld) && (res == nullptr || ld_ok(res_ld)))) &&
                  (!proj || (ld_ok(pa_ld) && ld_ok(pb_ld))),
              GTB_ERR_BAD_ARG, "gtb_in_node_fused_f32: row strides must cover 64 columns in 16-byte steps");
  if (n_nodes == 0) return GTB_OK;
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = nw_configure<true, true>();
    if (e == cudaSuccess) e = nw_configure<true, false>();
    if (e == cudaSuccess) e = nw_configure<false, true>();
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(in_node_ws)");
    configured = true;
  }
  NwParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.x_ld = x_ld; p.relu_x = relu_x;
  p.aggr = aggr; p.aggr_ld = aggr_ld; p.zero_aggr = zero_aggr;
  p.res = obj ? res : nullptr; p.res_ld = res_ld; p.res_a = res_a; p.res_b = res_b;
  p.x_out = x_out; p.xo_ld = xo_ld;
  p.pa = p_a; p.pa_ld = pa_ld; p.pb = p_b; p.pb_ld = pb_ld; p.proj_relu = proj_relu;
  p.packed_obj = static_cast<const unsigned char*>(packed_obj);
  p.packed_pa = static_cast<const unsigned char*>(packed_pa);
  p.packed_pb = static_cast<const unsigned char*>(packed_pb);
  p.n_rows = n_nodes;
  p.n_tiles = (int32_t)((n_nodes + NW_TM - 1) / NW_TM);
  const int pairs = (p.n_tiles + 1) / 2;
  const int grid = pairs < kNumSMs ? pairs : kNumSMs;
  if (obj && proj) in_node_ws_kernel<true, true><<<grid, NW_THREADS, NW_SMEM, st>>>(p);
  else if (obj)    in_node_ws_kernel<true, false><<<grid, NW_THREADS, NW_SMEM, st>>>(p);
  else             in_node_ws_kernel<false, true><<<grid, NW_THREADS, NW_SMEM, st>>>(p);
  GTB_CHECK_LAUNCH("in_node_ws_kernel");
  return GTB_OK;
}

int nw_fault_flag(int* out) { return check_cuda(cudaMemcpyFromSymbol(out, g_nw_fault, sizeof(int)), "cudaMemcpyFromSymbol(g_nw_fault)"); }

}  // namespace gtb
