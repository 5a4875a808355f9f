// W head of the edge classifier in one warp-specialised launch (sm_100a), the 64-wide "wide" shape:
//
//   w[o(r)] = eps + (1 - 2 eps) sigmoid(w2 . relu(W1 relu(sum_k W0_k e_k[i_k(r)] + T_dst[dst(r)] + T_src[src(r)] + b0) + b1) + b2)
//
// i.e. reference models/edge_classifier.py:108-117 -- the MLP `W` over cat[h[src], h[dst], e_0 .. e_3] -- with the
// two node blocks taken per node (T_src = h W0[:, 0:64]^T, T_dst = h W0[:, 64:128]^T, computed by the last node
// launch of the stack: GTB_SRC_PROJECTED) and the four 64-column edge-embedding blocks streamed as K chunks of
// the first Linear.  1 KB in, 4 bytes out per edge.
//
// The generic tiles (mlp_tc.cu) keep the 160 KB of hi / lo weights of this MLP resident, which leaves them ONE
// staging slot per team: every gather latency is exposed (435 us, 37 % of the HBM peak).  Here the first Linear's
// weights are STREAMED: the 32 KB hi / lo image of K chunk k travels from L2 into a two-buffer ring right beside
// the edge-embedding tiles it multiplies (one bulk copy per chunk and PAIR of tiles, ~1.7 MB per SM and launch),
// so four 32 KB data slots fit next to the resident second Linear.  Machine mapping as in edge_ws.cu:
//
//   * one persistent CTA per SM, two tiles ("contexts") in flight; warps 0-15: 512 row owners (thread = row x
//     16-column quarter) alternating between the contexts chunk by chunk; warps 16 / 17: TMA producers (edge tiles by
//     2-D tile loads or tile::gather4 through `perm`, T_src rows by gather4, ids, and -- warp 16 -- the weight
//     chunks); warps 18 / 19: tcgen05.mma issue (24 MMAs per chunk / Linear, A operand in TMEM, 3xTF32);
//   * per tile five staged items (e_0 .. e_3, T_src rows) go through the context's two slots, six accumulator
//     hand-overs (four chunks accumulate into one D, then the second Linear); the last Linear (64 -> 1) runs on
//     the CUDA cores: 16 FMAs per thread, the four quarter sums of a row meet in spare TMEM columns;
//   * every hand-over is an mbarrier: full / free per slot, a_ready / d_ready, red_ready, wfull / wfree per
//     weight buffer (wfree counts the commits of BOTH contexts' MMA warps).
#include "common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace gtb {

using namespace tc;

size_t tc_packed_bytes(int, const int32_t*, int, const int32_t*);

constexpr int HW_TM = 128;
constexpr int HW_W1 = 0;                 // second Linear: hi 16 KB | lo 16 KB
constexpr int HW_RING = 32768;           // two buffers of (hi 16 KB | lo 16 KB): K chunk k of the first Linear in buffer k & 1
constexpr int HW_SLOTS = 98304;          // context c, slot s at HW_SLOTS + (2 c + s) * 32768
constexpr int HW_IDS = 229376;           // per context: dst ids [128] | src ids [128]
constexpr int HW_BIAS = HW_IDS + 2048;   // b0 [64] | b1 [64] | w2 [64] | b2
constexpr int HW_BARS = HW_BIAS + 784;   // per context 64 bytes: full[2] | free[2] | a_ready | d_ready | red_ready; then wfull[2] | wfree[2]
constexpr int HW_TMEM_SLOT = HW_BARS + 160;
constexpr int HW_SMEM = HW_TMEM_SLOT + 8;
constexpr int HW_THREADS = 640;
constexpr uint32_t HW_A_HI = 0, HW_A_LO = 64, HW_D = 128, HW_CTX = 192, HW_RED = 384;
// the gtb_mlp_pack image of dims {256, 64, 64, 1} with blocks {64, 64, 64, 64} (tc_layout in mlp_tc.cu)
constexpr int HP_W0_HI = 0, HP_W0_LO = 65536, HP_W1 = 131072, HP_W2 = 163840, HP_B0 = 164096, HP_B1 = 164352, HP_B2 = 164608;
constexpr size_t HP_BYTES = 164624;
static_assert(HW_SMEM <= 232448, "shared-memory layout");

__device__ int g_hw_fault = 0;

struct HwParams {
  CUtensorMap e_map[4];  // e_k [*, 64] fp32: box 32 x 128 (tile mode) or 32 x 1 (gather mode)
  CUtensorMap ts_map;    // T_src [*, 64] fp32: box 32 x 1
  const int32_t* e_index[4];
  const float* tdst;
  const int32_t* dst;
  const int32_t* src;
  const int32_t* out_index;
  float* out;
  const unsigned char* packed;
  int64_t n_rows;
  int32_t n_tiles, tdst_ld, out_ld, final_act;
  float act_eps;
  float* h0;  // optional: post-ReLU outputs of the first / second Linear, [n_rows, h_ld] in launch-row order
  float* h1;
  int32_t h_ld;
};

__device__ __noinline__ void hw_timeout() {
  atomicExch(&g_hw_fault, 1);
  __trap();
}
__device__ __forceinline__ void hw_wait(uint32_t bar, uint32_t parity) {
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
  hw_timeout();
}

// use number (0-based, over the whole kernel) of slot s = j & 1 of a context by item j of tile iteration t:
// slot 0 takes items 0, 2, 4 (three per tile), slot 1 items 1, 3
__device__ __forceinline__ uint32_t hw_use(int t, int j) { return (j & 1) ? 2u * t + (j >> 1) : 3u * t + (j >> 1); }

// row indices 4 lane .. 4 lane + 3 of a tile (`fill` past the end of the list)
__device__ __forceinline__ int4 hw_idx4(const int32_t* idx, uint32_t row0, int rows_here, int lane, int fill) {
  if (rows_here == HW_TM) return __ldg(reinterpret_cast<const int4*>(idx + row0) + lane);
  int4 v;
  v.x = 4 * lane + 0 < rows_here ? __ldg(idx + row0 + 4 * lane + 0) : fill;
  v.y = 4 * lane + 1 < rows_here ? __ldg(idx + row0 + 4 * lane + 1) : fill;
  v.z = 4 * lane + 2 < rows_here ? __ldg(idx + row0 + 4 * lane + 2) : fill;
  v.w = 4 * lane + 3 < rows_here ? __ldg(idx + row0 + 4 * lane + 3) : fill;
  return v;
}

struct HwOwner {
  const HwParams& p;
  uint32_t sm0;
  int w, lane, r, qd;
  uint32_t rx, own;
  int tile00, nA, nB;

  __device__ __forceinline__ int n_of(int c) const { return c ? nB : nA; }
  __device__ __forceinline__ int tile_of(int c, int t) const { return tile00 + (2 * t + c) * (int)gridDim.x; }
  __device__ __forceinline__ uint32_t bar(int c, int which) const { return sm0 + HW_BARS + 64 * c + 8 * which; }
  __device__ __forceinline__ uint32_t slot(int c, int s) const { return sm0 + HW_SLOTS + (2 * c + s) * 32768; }
  __device__ __forceinline__ uint32_t tm_lane(int c) const { return (uint32_t)c * HW_CTX + ((uint32_t)((w & 3) * 32) << 16); }
  __device__ __forceinline__ uint32_t chunk(int q) const { return (((uint32_t)(4 * (qd & 1) + q)) << 4) ^ rx; }
  __device__ __forceinline__ void a_done(int c) {
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) tma::mbar_arrive(bar(c, 4));
  }
  // completion number g (0 .. 3: K chunks, 4: second Linear) of tile iteration t
  __device__ __forceinline__ void wait_d(int c, int t, int g) {
    hw_wait(bar(c, 5), (uint32_t)(5 * t + g) & 1u);
    tc_fence_after_sync();
  }
  __device__ __forceinline__ void acc_load(int c, float (&v)[16]) {
    uint32_t acc[16];
    tmem_ld16(tm_lane(c) + HW_D + 16 * qd, acc);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
  }
  __device__ __forceinline__ void bias_add(float (&v)[16], int layer) {
    const uint32_t ba = sm0 + HW_BIAS + 256 * layer + 64 * qd;
    add16(v, lds128(ba), lds128(ba + 16), lds128(ba + 32), lds128(ba + 48));
  }

  __device__ __forceinline__ void save16(float* table, int c, int t, const float (&v)[16]) const {
    const int64_t row = (int64_t)tile_of(c, t) * HW_TM + r;
    if (row < p.n_rows) {
      float4* q = reinterpret_cast<float4*>(table + row * (int64_t)p.h_ld + 16 * qd);
#pragma unroll
      for (int i = 0; i < 4; ++i) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
  }

  // ---- C0: K chunk k (edge-embedding block k): own 16 columns -> tf32 hi / lo -> TMEM
  __device__ __forceinline__ void c0(int c, int t, int k) {
    const int s = k & 1;
    hw_wait(bar(c, s), hw_use(t, k) & 1u);
    const uint32_t sl = slot(c, s);
    float v[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = lds128(sl + own + chunk(q));
      v[4 * q + 0] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
    }
    uint32_t hi[16], lo[16];
    split16(v, hi, lo);              // in front of the wait: only the two TMEM stores are left behind it
    if (k > 0) wait_d(c, t, k - 1);  // the previous chunk's MMAs have read the A columns
    tmem_st16(tm_lane(c) + HW_A_HI + 16 * qd, hi);
    tmem_st16(tm_lane(c) + HW_A_LO + 16 * qd, lo);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) {
      tma::mbar_arrive(bar(c, 4));      // a_ready
      tma::mbar_arrive(bar(c, 2 + s));  // the slot is free (its piece sits in registers / TMEM)
    }
  }

  // ---- E0: D + b0 + T_dst[dst] + T_src[src] -> ReLU -> A operand of the second Linear
  __device__ __forceinline__ void e0(int c, int t) {
    const uint32_t row0 = (uint32_t)tile_of(c, t) * HW_TM;
    const int rows_here = (int)min((int64_t)HW_TM, p.n_rows - (int64_t)row0);
    uint4 pre[4];
    {
      int32_t d = 0;
      if (rows_here == HW_TM) d = lds_i32(sm0 + HW_IDS + 1024 * c + 4 * r);
      else if (r < rows_here) d = __ldg(p.dst + row0 + r);
      const uint4* rowp = reinterpret_cast<const uint4*>(p.tdst + (uint64_t)(uint32_t)d * (uint32_t)p.tdst_ld + 16 * qd);
#pragma unroll
      for (int q = 0; q < 4; ++q) pre[q] = __ldg(rowp + q);
    }
    {  // T_src rows (item 4, slot 0) and the last K chunk's MMAs: both first tries in front of the first branch
      const uint32_t ba = bar(c, 0), pa = hw_use(t, 4) & 1u, bb = bar(c, 5), pb = (uint32_t)(5 * t + 3) & 1u;
      uint32_t ok_a, ok_b;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok_a) : "r"(ba), "r"(pa) : "memory");
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok_b) : "r"(bb), "r"(pb) : "memory");
      if (!(ok_a & ok_b)) {
        if (!ok_a) hw_wait(ba, pa);
        if (!ok_b) hw_wait(bb, pb);
      }
      tc_fence_after_sync();
    }
    float v[16];
    acc_load(c, v);
    const uint32_t sl = slot(c, 0);
    float4 x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) x[q] = lds128(sl + own + chunk(q));
    bias_add(v, 0);
    add16(v, make_float4(__uint_as_float(pre[0].x), __uint_as_float(pre[0].y), __uint_as_float(pre[0].z), __uint_as_float(pre[0].w)),
          make_float4(__uint_as_float(pre[1].x), __uint_as_float(pre[1].y), __uint_as_float(pre[1].z), __uint_as_float(pre[1].w)),
          make_float4(__uint_as_float(pre[2].x), __uint_as_float(pre[2].y), __uint_as_float(pre[2].z), __uint_as_float(pre[2].w)),
          make_float4(__uint_as_float(pre[3].x), __uint_as_float(pre[3].y), __uint_as_float(pre[3].z), __uint_as_float(pre[3].w)));
    add16(v, x[0], x[1], x[2], x[3]);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
    split_store16(tm_lane(c) + HW_A_HI + 16 * qd, tm_lane(c) + HW_A_LO + 16 * qd, v);
    if (p.h0 != nullptr) save16(p.h0, c, t, v);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) {
      tma::mbar_arrive(bar(c, 4));
      tma::mbar_arrive(bar(c, 2));
    }
  }

  // ---- E1: D + b1 -> ReLU -> own quarter of the dot product with w2 -> spare TMEM column
  __device__ __forceinline__ void e1(int c, int t) {
    wait_d(c, t, 4);
    float v[16];
    acc_load(c, v);
    bias_add(v, 1);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
    if (p.h1 != nullptr) save16(p.h1, c, t, v);
    const uint32_t wa = sm0 + HW_BIAS + 512 + 64 * qd;
    float part = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 w4 = lds128(wa + 16 * q);
      part = fmaf(v[4 * q + 0], w4.x, part);
      part = fmaf(v[4 * q + 1], w4.y, part);
      part = fmaf(v[4 * q + 2], w4.z, part);
      part = fmaf(v[4 * q + 3], w4.w, part);
    }
    tmem_st1(tm_lane(c) - (uint32_t)c * HW_CTX + HW_RED + 4 * c + qd, __float_as_uint(part));
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) tma::mbar_arrive(bar(c, 6));
  }

  // ---- FIN (quarter 0 threads: one per row): the four quarter sums + b2 -> activation -> store
  __device__ __forceinline__ void fin(int c, int t) {
    if (qd != 0) return;
    const uint32_t row0 = (uint32_t)tile_of(c, t) * HW_TM;
    const int rows_here = (int)min((int64_t)HW_TM, p.n_rows - (int64_t)row0);
    int64_t orow = (int64_t)row0 + r;
    if (p.out_index != nullptr && r < rows_here) orow = __ldg(p.out_index + row0 + r);
    hw_wait(bar(c, 6), (uint32_t)t & 1u);
    tc_fence_after_sync();
    uint32_t q4[4];
    tmem_ld4(tm_lane(c) - (uint32_t)c * HW_CTX + HW_RED + 4 * c, q4);
    tmem_ld_wait();
    float v = (__uint_as_float(q4[0]) + __uint_as_float(q4[1])) + (__uint_as_float(q4[2]) + __uint_as_float(q4[3]));
    v += lds32(sm0 + HW_BIAS + 768);
    if (p.final_act == GTB_ACT_RELU) v = fmaxf(v, 0.f);
    else if (p.final_act == GTB_ACT_SIGMOID_AFFINE) v = p.act_eps + (1.f - 2.f * p.act_eps) * (1.f / (1.f + expf(-v)));
    if (r < rows_here) p.out[orow * (int64_t)p.out_ld] = v;
    tc_fence_before_sync();
  }
};

__global__ void __launch_bounds__(HW_THREADS, 1) ec_head_ws_kernel(const __grid_constant__ HwParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sm0 = smem_u32(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + HW_TMEM_SLOT);
  if (sm0 & 1023u) {
    if (tid == 0) atomicExch(&g_hw_fault, 3);
    __trap();
  }
  {  // resident pieces: second Linear (hi | lo), biases, w2 (generic proxy; visible to the tensor core after the proxy fence)
    const float4* g4 = reinterpret_cast<const float4*>(p.packed + HP_W1);
    float4* s4 = reinterpret_cast<float4*>(smem_raw + HW_W1);
    for (int i = tid; i < 32768 / 16; i += HW_THREADS) s4[i] = __ldg(g4 + i);
    float* sb = reinterpret_cast<float*>(smem_raw + HW_BIAS);
    if (tid < 64) {
      sb[tid] = __ldg(reinterpret_cast<const float*>(p.packed + HP_B0) + tid);
      sb[64 + tid] = __ldg(reinterpret_cast<const float*>(p.packed + HP_B1) + tid);
      sb[128 + tid] = __ldg(reinterpret_cast<const float*>(p.packed + HP_W2) + tid);
    }
    if (tid == 64) sb[192] = __ldg(reinterpret_cast<const float*>(p.packed + HP_B2));
  }
  if (tid == 0) {
    for (int c = 0; c < 2; ++c) {
      unsigned char* b = smem_raw + HW_BARS + 64 * c;
      mbar_init(reinterpret_cast<uint64_t*>(b + 0), 1);    // full[0]
      mbar_init(reinterpret_cast<uint64_t*>(b + 8), 1);    // full[1]
      mbar_init(reinterpret_cast<uint64_t*>(b + 16), 16);  // free[0]
      mbar_init(reinterpret_cast<uint64_t*>(b + 24), 16);  // free[1]
      mbar_init(reinterpret_cast<uint64_t*>(b + 32), 16);  // a_ready
      mbar_init(reinterpret_cast<uint64_t*>(b + 40), 1);   // d_ready
      mbar_init(reinterpret_cast<uint64_t*>(b + 48), 16);  // red_ready
    }
    unsigned char* wb = smem_raw + HW_BARS + 128;
    mbar_init(reinterpret_cast<uint64_t*>(wb + 0), 1);   // wfull[0]
    mbar_init(reinterpret_cast<uint64_t*>(wb + 8), 1);   // wfull[1]
    mbar_init(reinterpret_cast<uint64_t*>(wb + 16), 2);  // wfree[0]: both contexts' MMA warps
    mbar_init(reinterpret_cast<uint64_t*>(wb + 24), 2);  // wfree[1]
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (*tmem_slot != 0) {
    if (tid == 0) atomicExch(&g_hw_fault, 2);
    __trap();
  }

  const int g = (int)gridDim.x;
  int tile0[2], n_t[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    tile0[c] = (int)blockIdx.x + g * c;
    n_t[c] = tile0[c] < p.n_tiles ? (p.n_tiles - tile0[c] + 2 * g - 1) / (2 * g) : 0;
  }
  const uint32_t wfull = sm0 + HW_BARS + 128, wfree = wfull + 16;

  if (warp < 16) {
    // ================================================================= row owners
    HwOwner o{p, sm0};
    o.w = warp; o.lane = lane; o.r = 32 * (warp & 3) + lane; o.qd = warp >> 2;
    o.rx = (uint32_t)(o.r & 7) << 4;
    o.own = (uint32_t)(o.qd >> 1) * 16384u + (uint32_t)o.r * 128u;
    o.tile00 = tile0[0]; o.nA = n_t[0]; o.nB = n_t[1];
#pragma unroll 1
    for (int t = 0; t < o.nA; ++t) {
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
#pragma unroll 1
        for (int c = 0; c < 2; ++c)
          if (t < o.n_of(c)) o.c0(c, t, k);
      }
#pragma unroll 1
      for (int c = 0; c < 2; ++c)
        if (t < o.n_of(c)) o.e0(c, t);
#pragma unroll 1
      for (int c = 0; c < 2; ++c)
        if (t < o.n_of(c)) o.e1(c, t);
#pragma unroll 1
      for (int c = 0; c < 2; ++c)
        if (t < o.n_of(c)) o.fin(c, t);
    }
  } else if (warp < 18) {
    // ================================================================= TMA producer of context `ctx` (warp 16 also streams the weights)
    const int ctx = warp - 16;
    const uint32_t bars = sm0 + HW_BARS + 64 * ctx;
    const uint32_t ids = sm0 + HW_IDS + 1024 * ctx;
    if (lane == 0) {
      for (int k = 0; k < 4; ++k) tma::prefetch_map(&p.e_map[k]);
      tma::prefetch_map(&p.ts_map);
    }
    const int n_rounds = n_t[0];  // >= n_t[1]
    for (int t = 0; t < n_rounds; ++t) {
      const bool have = t < n_t[ctx];
      const int tile = tile0[ctx] + t * 2 * g;
      const uint32_t row0 = (uint32_t)tile * HW_TM;
      const int rows_here = have ? (int)min((int64_t)HW_TM, p.n_rows - (int64_t)row0) : 0;
      const bool full_tile = rows_here == HW_TM;
#pragma unroll 1
      for (int j = 0; j < 5; ++j) {
        if (ctx == 0 && j < 4) {  // K chunk j of the first Linear for this round's pair of tiles
          const int b = j & 1;
          const uint32_t u = 2u * t + (j >> 1);
          if (u >= 1) hw_wait(wfree + 8 * b, (u - 1) & 1u);
          if (elect_one()) {
            tma::mbar_expect_tx(wfull + 8 * b, 32768);
            tma::bulk_g2s(sm0 + HW_RING + b * 32768, p.packed + HP_W0_HI + j * 16384, 16384, wfull + 8 * b);
            tma::bulk_g2s(sm0 + HW_RING + b * 32768 + 16384, p.packed + HP_W0_LO + j * 16384, 16384, wfull + 8 * b);
          }
          __syncwarp();
        }
        if (!have) continue;
        const int s = j & 1;
        const uint32_t u = hw_use(t, j);
        const uint32_t full = bars + 8 * s, free_ = bars + 16 + 8 * s;
        const uint32_t sl = sm0 + HW_SLOTS + (2 * ctx + s) * 32768;
        if (u >= 1) hw_wait(free_, (u - 1) & 1u);
        const bool with_ids = j == 0 && full_tile;
        if (lane == 0) tma::mbar_expect_tx(full, 32768 + (with_ids ? 1024 : 0));
        __syncwarp();
        const int32_t* index = j < 4 ? p.e_index[j] : p.src;
        const CUtensorMap* map = j < 4 ? &p.e_map[j] : &p.ts_map;
        if (index != nullptr) {
          const int4 r4 = hw_idx4(index, row0, rows_here, lane, 0);
#pragma unroll 4
          for (int i = 0; i < 32; ++i) {
            const int a = __shfl_sync(0xffffffffu, r4.x, i), b = __shfl_sync(0xffffffffu, r4.y, i);
            const int cc = __shfl_sync(0xffffffffu, r4.z, i), d = __shfl_sync(0xffffffffu, r4.w, i);
            if (elect_one()) {
              tma::gather4(sl + i * 512, map, full, 0, a, b, cc, d);
              tma::gather4(sl + 16384 + i * 512, map, full, 32, a, b, cc, d);
            }
            __syncwarp();
          }
        }
        if (elect_one()) {
          if (index == nullptr) {
            tma::load_2d(sl, map, full, 0, (int)row0);  // rows past the end of the table arrive as zeros
            tma::load_2d(sl + 16384, map, full, 32, (int)row0);
          }
          if (with_ids) {
            tma::bulk_g2s(ids, p.dst + row0, 512, full);
            tma::bulk_g2s(ids + 512, p.src + row0, 512, full);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================= MMA issue of context `ctx`
    const int ctx = warp - 18;
    const uint32_t a_ready = sm0 + HW_BARS + 64 * ctx + 32, d_ready = a_ready + 8;
    const uint32_t tmc = (uint32_t)ctx * HW_CTX;
    const uint32_t idesc = make_idesc_tf32(HW_TM, 64);
    const int n_rounds = n_t[0];
    int n = 0;  // a_ready completions consumed
    for (int t = 0; t < n_rounds; ++t) {
      const bool have = t < n_t[ctx];
#pragma unroll 1
      for (int l = 0; l < 5; ++l) {
        uint32_t w_hi = sm0 + HW_W1, w_lo = sm0 + HW_W1 + 16384u;
        if (l < 4) {
          const int b = l & 1;
          w_hi = sm0 + HW_RING + b * 32768;
          w_lo = w_hi + 16384u;
          hw_wait(wfull + 8 * b, (2u * t + (l >> 1)) & 1u);  // also paces a context without a tile this round
          if (!have) {
            if (lane == 0) tma::mbar_arrive(wfree + 8 * b);
            __syncwarp();
            continue;
          }
        } else if (!have) {
          continue;
        }
        hw_wait(a_ready, (uint32_t)n & 1u);
        ++n;
        tc_fence_after_sync();
        const uint64_t bd_hi = make_smem_desc_sw128(w_hi), bd_lo = make_smem_desc_sw128(w_lo);
        if (elect_one()) {
          bool acc = l > 0 && l < 4;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {  // small terms first: lo*hi, hi*lo, hi*hi
            const uint32_t a = tmc + ((pass == 0) ? HW_A_LO : HW_A_HI);
            const uint64_t bd = (pass == 1) ? bd_lo : bd_hi;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              mma_tf32_ts(tmc + HW_D, a + 8 * ks, bd + (uint64_t)((ks >> 2) * 512 + (ks & 3) * 2), idesc, acc);
              acc = true;
            }
          }
          mma_commit_addr(d_ready);
          if (l < 4) mma_commit_addr(wfree + 8 * (l & 1));
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0u, 512);
}

// Does this descriptor have the shape of the wide W head?  Returns the positions of the source blocks.
static bool hw_match(const gtb_mlp_desc_t& d, int* s_src, int* s_dst, int (&s_e)[4]) {
  if (d.n_layers != 3 || d.n_srcs != 6) return false;
  if (d.dims[0] != 256 || d.dims[1] != 64 || d.dims[2] != 64 || d.dims[3] != 1) return false;
  if (d.out == nullptr || d.aggr != nullptr || d.res != nullptr || d.res_b != 1.f || d.row_scale || d.out_scale || d.gate) return false;
  if (d.final_act != GTB_ACT_SIGMOID_AFFINE && d.final_act != GTB_ACT_NONE && d.final_act != GTB_ACT_RELU) return false;
  *s_src = *s_dst = -1;
  int ne = 0;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  for (int s = 0; s < 6; ++s) {
    const gtb_src_t& b = d.srcs[s];
    if (b.width != 64 || (b.ld & 3) || !al16(b.ptr) || b.relu) return false;
    if (b.index != nullptr && !al16(b.index)) return false;
    if (b.flags & GTB_SRC_PROJECTED) {
      if (b.index == nullptr) return false;
      if ((b.flags & GTB_SRC_SORTED) && *s_dst < 0) *s_dst = s;  // gathered by the sorted destination ids
      else if (*s_src < 0) *s_src = s;
      else return false;
    } else {
      if (ne == 4) return false;
      s_e[ne++] = s;
    }
  }
  if (ne != 4 || *s_src < 0 || *s_dst < 0) return false;
  if (d.out_index != nullptr && !al16(d.out_index)) return false;
  return true;
}

int ec_head_ws(const gtb_mlp_desc_t& d, cudaStream_t st, bool* handled) {
  *handled = false;
  static const bool disabled = getenv("GTB_NO_HEAD_WS") != nullptr;
  int s_src, s_dst, s_e[4];
  if (disabled || !hw_match(d, &s_src, &s_dst, s_e) || tma::encode_fn() == nullptr) return GTB_OK;
  {  // the packed image must be the one the offsets above describe
    const int32_t dims[4] = {256, 64, 64, 1}, bw[4] = {64, 64, 64, 64};
    if (tc_packed_bytes(3, dims, 4, bw) != HP_BYTES) return GTB_OK;
  }
  HwParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t big = 1ull << 31;  // row bound of gathered tables (indices come from a validated plan)
  for (int k = 0; k < 4; ++k) {
    const gtb_src_t& b = d.srcs[s_e[k]];
    const bool gather = b.index != nullptr;
    if (!tma::make_map_2d(&p.e_map[k], b.ptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, gather ? big : (uint64_t)d.n_rows, 64, (uint64_t)b.ld, 32,
                          gather ? 1 : 128))
      return GTB_OK;  // the driver refused a map: the generic tiles take the launch
    p.e_index[k] = b.index;
  }
  const gtb_src_t &bs = d.srcs[s_src], &bd = d.srcs[s_dst];
  if (!tma::make_map_2d(&p.ts_map, bs.ptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, big, 64, (uint64_t)bs.ld, 32, 1)) return GTB_OK;
  p.tdst = bd.ptr;
  p.tdst_ld = bd.ld;
  p.dst = bd.index;
  p.src = bs.index;
  p.out_index = d.out_index;
  p.out = d.out;
  p.out_ld = d.out_ld;
  p.packed = static_cast<const unsigned char*>(d.packed);
  p.n_rows = d.n_rows;
  p.n_tiles = (int32_t)((d.n_rows + HW_TM - 1) / HW_TM);
  p.final_act = d.final_act;
  p.act_eps = d.act_eps;
  if (d.hidden0 != nullptr || d.hidden1 != nullptr) {
    GTB_REQUIRE(d.hidden0 != nullptr && d.hidden1 != nullptr && d.hidden_ld >= 64 && !(d.hidden_ld & 3) &&
                    !(reinterpret_cast<uintptr_t>(d.hidden0) & 15) && !(reinterpret_cast<uintptr_t>(d.hidden1) & 15),
                GTB_ERR_BAD_ARG, "gtb_fused_mlp_f32: hidden0 / hidden1 come as a pair of 16-byte aligned [n_rows, >= 64] tables");
    p.h0 = d.hidden0;
    p.h1 = d.hidden1;
    p.h_ld = d.hidden_ld;
  }
  *handled = true;
  if (d.n_rows == 0) return GTB_OK;
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ec_head_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HW_SMEM);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(ec_head_ws)");
    configured = true;
  }
  const int pairs = (p.n_tiles + 1) / 2;
  const int grid = pairs < kNumSMs ? pairs : kNumSMs;
  ec_head_ws_kernel<<<grid, HW_THREADS, HW_SMEM, st>>>(p);
  GTB_CHECK_LAUNCH("ec_head_ws_kernel");
  return GTB_OK;
}

bool ec_head_ws_takes(const gtb_mlp_desc_t& d) {
  static const bool disabled = getenv("GTB_NO_HEAD_WS") != nullptr;
  int s_src, s_dst, s_e[4];
  const int32_t dims[4] = {256, 64, 64, 1}, bw[4] = {64, 64, 64, 64};
  return !disabled && hw_match(d, &s_src, &s_dst, s_e) && tma::encode_fn() != nullptr && tc_packed_bytes(3, dims, 4, bw) == HP_BYTES;
}

int hw_fault_flag(int* out) { return check_cuda(cudaMemcpyFromSymbol(out, g_hw_fault, sizeof(int)), "cudaMemcpyFromSymbol(g_hw_fault)"); }

}  // namespace gtb
