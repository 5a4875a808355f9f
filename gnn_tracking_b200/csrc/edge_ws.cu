// Warp-specialised fused Interaction-Network edge kernel (sm_100a), the 64 / 64 / 64 "wide" shape:
//
//   e_out[o(r)] = W2 relu(W1 relu(W0e e_in[i(r)] + P_i[dst(r)] + P_j[src(r)] + b0) + b1) + b2
//   aggr[dst(r)] += e_out row          (rows r walk the plan's destination-sorted edge list)
//
// i.e. reference models/interaction_network.py:75-89 (message) + the SumAggregation of :22,36 with the
// node-side products P_i = relu(x) W0[:, :Dn]^T, P_j = relu(x) W0[:, Dn:2Dn]^T taken per node (see
// GTB_SRC_PROJECTED in include/gtb200.h).  Same arithmetic as the generic tiles of mlp_tc.cu (3xTF32,
// A operand in TMEM, fp32 accumulation), different machine mapping:
//
//   * one persistent CTA per SM, 20 warps with fixed roles, two tile contexts (ctx = alternate tiles):
//       warp 0 / 1  : TMA producer of ctx 0 / 1 -- edge-feature tile (2-D tile load when the features are
//                     kept in destination order, tile::gather4 through `perm` otherwise), P_j rows
//                     (tile::gather4 through src_sorted), and the store of the finished tile
//                     (tile store / tile::scatter4); completion through mbarrier transaction counts
//       warp 2 / 3  : tcgen05.mma issue of ctx 0 / 1 (one elected lane, 24 MMAs per Linear)
//       warps 4-11 / 12-19 : 256 row-owner threads of ctx 0 / 1 (thread = row x 32-column half):
//                     tf32 hi / lo split into TMEM, accumulator read-back, bias / gathered adds / ReLU,
//                     output staging and the in-tile segmented sum
//   * every hand-over is an mbarrier (no CTA or named barrier inside the tile loop):
//       full[slot]  producer -> owners   (TMA bytes landed)
//       empty[slot] owners   -> producer (slot may be overwritten)
//       a_ready     owners   -> MMA warp (A operand of the next Linear is in TMEM)
//       d_ready     MMA warp -> owners   (tcgen05.commit: accumulator complete, A operand free)
//       out_ready   owners   -> producer + owners (output tile staged in shared memory)
//   * shared memory: packed weights (3 x hi / lo x 16 KB, the image gtb_mlp_pack writes) + a ring of two
//     32 KB slots per context; per tile the ring carries e_in, P_j, out in that order, so e_in(t + 1)
//     lands while tile t is still in its second Linear.
//   * TMEM: 192 columns per context (A hi | A lo | D).
#include "common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace gtb {

using namespace tc;

constexpr int EW_TM = 128;
constexpr int EW_SLOT = 32768;                // [2 K tiles][128 rows][32 fp32], 128-byte swizzle (TMA == UMMA image)
constexpr int EW_HEAD = 2048;                 // mbarriers + TMEM slot
constexpr int EW_WBYTES = 99328;              // packed weights (99072 bytes) rounded up to 1024
constexpr int EW_SMEM = EW_HEAD + EW_WBYTES + 4 * EW_SLOT;  // 232448 = the 227 KB opt-in maximum
constexpr int EW_THREADS = 640;
constexpr uint32_t EW_A_HI = 0, EW_A_LO = 64, EW_D = 128, EW_CTX = 192;

__device__ int g_ew_fault = 0;  // 1: barrier timeout, 2: TMEM base != 0, 3: shared memory misaligned
__device__ long long g_ew_prof[32];
static int g_ew_prof_enabled = 0;

// per-stage clock accumulation by lane 0 of one warp per role of context 0 in CTA 0 (tests/cuda/tc_diag.py)
#define EW_PROF(id)                                   \
  do {                                                \
    if (PROF && prof_on) {                            \
      const long long now_ = clock64();               \
      g_ew_prof[id] += now_ - prof_t;                 \
      prof_t = now_;                                  \
    }                                                 \
  } while (0)

struct EwParams {
  CUtensorMap e_map;    // e_in  [E, 64] fp32: box 32 x 128 (tile mode) or 32 x 1 (gather mode)
  CUtensorMap pj_map;   // P_j   [N, 64] fp32: box 32 x 1
  CUtensorMap out_map;  // e_out [E, 64] fp32: box 32 x 128 (tile mode) or 32 x 1 (scatter mode)
  const float* pi;      // P_i [N, pi_ld]
  const int32_t* dst;   // dst_sorted [E]: segment ids of the per-destination sum
  const int32_t* pi_index;  // row of P_i per edge (the same array in an IN layer)
  const int32_t* src;   // src_sorted [E]
  const int32_t* e_index;    // perm or nullptr (tile mode)
  const int32_t* out_index;  // perm or nullptr (tile mode)
  float* aggr;
  float* out;           // e_out base (scatter mode: the rows of a partial last tile are stored by their owners)
  const unsigned char* packed;
  int64_t n_rows;
  int32_t n_tiles, pi_ld, aggr_ld, out_ld, w_bytes;
};

__device__ __forceinline__ void ew_wait(uint32_t bar, uint32_t parity) {
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
  atomicExch(&g_ew_fault, 1);
  __trap();  // a protocol error must fail loudly, never hang the GPU
}

// row indices 4 lane .. 4 lane + 3 of a tile (clamped / replaced by `fill` past the end of the list)
__device__ __forceinline__ int4 ew_idx4(const int32_t* idx, uint32_t row0, int rows_here, int lane, int fill) {
  if (rows_here == EW_TM) return __ldg(reinterpret_cast<const int4*>(idx + row0) + lane);
  int4 v;
  v.x = 4 * lane + 0 < rows_here ? __ldg(idx + row0 + 4 * lane + 0) : fill;
  v.y = 4 * lane + 1 < rows_here ? __ldg(idx + row0 + 4 * lane + 1) : fill;
  v.z = 4 * lane + 2 < rows_here ? __ldg(idx + row0 + 4 * lane + 2) : fill;
  v.w = 4 * lane + 3 < rows_here ? __ldg(idx + row0 + 4 * lane + 3) : fill;
  return v;
}

// 64 gather4 copies of one 128-row tile: lane j's four row indices are broadcast, ONE elected lane issues
// (uniform-datapath instruction: coordinates travel through uniform registers, no per-lane waterfall)
__device__ __forceinline__ void ew_gather_tile(uint32_t slot, const CUtensorMap* map, uint32_t bar, const int4& r4) {
#pragma unroll 4
  for (int j = 0; j < 32; ++j) {
    const int a = __shfl_sync(0xffffffffu, r4.x, j), b = __shfl_sync(0xffffffffu, r4.y, j);
    const int c = __shfl_sync(0xffffffffu, r4.z, j), d = __shfl_sync(0xffffffffu, r4.w, j);
    if (elect_one()) {
      tma::gather4(slot + j * 512, map, bar, 0, a, b, c, d);
      tma::gather4(slot + 16384 + j * 512, map, bar, 32, a, b, c, d);
    }
    __syncwarp();
  }
}
__device__ __forceinline__ void ew_scatter_tile(uint32_t slot, const CUtensorMap* map, const int4& r4) {
#pragma unroll 4
  for (int j = 0; j < 32; ++j) {
    const int a = __shfl_sync(0xffffffffu, r4.x, j), b = __shfl_sync(0xffffffffu, r4.y, j);
    const int c = __shfl_sync(0xffffffffu, r4.z, j), d = __shfl_sync(0xffffffffu, r4.w, j);
    if (elect_one()) {
      tma::scatter4(map, slot + j * 512, 0, a, b, c, d);
      tma::scatter4(map, slot + 16384 + j * 512, 32, a, b, c, d);
    }
    __syncwarp();
  }
}

template <bool RELU_E, bool PROF>
__global__ void __launch_bounds__(EW_THREADS, 1) in_edge_ws_kernel(const __grid_constant__ EwParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sm0 = smem_u32(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // head: per context c (64 bytes each): full[2] | empty[2] | a_ready | d_ready | out_ready ; TMEM slot at 256
  const uint32_t wbase = sm0 + EW_HEAD;
  const uint32_t slots0 = wbase + EW_WBYTES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 256);

  if (sm0 & 1023u) {  // swizzle atoms need the 1024-byte alignment the layout above assumes
    if (tid == 0) atomicExch(&g_ew_fault, 3);
    __trap();
  }
  {  // packed weights -> shared memory (generic proxy), visible to the tensor core after the proxy fence
    const float4* g4 = reinterpret_cast<const float4*>(p.packed);
    float4* s4 = reinterpret_cast<float4*>(smem_raw + EW_HEAD);
    for (int i = tid; i < (p.w_bytes >> 4); i += EW_THREADS) s4[i] = __ldg(g4 + i);
  }
  if (tid == 0) {
    for (int c = 0; c < 2; ++c) {
      const uint32_t b = sm0 + 64 * c;
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + 64 * c + 0), 1);   // full[0]
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + 64 * c + 8), 1);   // full[1]
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + 64 * c + 16), 8);  // empty[0]: 8 owner warps
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + 64 * c + 24), 8);  // empty[1]
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + 64 * c + 32), 8);  // a_ready
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + 64 * c + 40), 1);  // d_ready
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + 64 * c + 48), 8);  // out_ready
      (void)b;
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (*tmem_slot != 0) {  // one CTA per SM, one allocation: column 0 / lane 0; the MMA issue relies on it
    if (tid == 0) atomicExch(&g_ew_fault, 2);
    __trap();
  }

  const int ctx = warp < 4 ? (warp & 1) : ((warp - 4) >> 3);
  const uint32_t bars = sm0 + 64 * ctx;
  const uint32_t full0 = bars, empty0 = bars + 16, a_ready = bars + 32, d_ready = bars + 40, out_ready = bars + 48;
  const uint32_t slots = slots0 + (uint32_t)ctx * 2u * EW_SLOT;
  const uint32_t tmc = (uint32_t)ctx * EW_CTX;
  // tiles of this context: blockIdx.x + gridDim.x * (2 t + ctx)
  const int tile0 = (int)blockIdx.x + (int)gridDim.x * ctx, tstep = 2 * (int)gridDim.x;
  const bool prof_on = PROF && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 2 || warp == 4);
  long long prof_t = prof_on ? clock64() : 0;

  if (warp < 2) {
    // ================================================================= TMA producer of context `ctx`
    if (lane == 0) {
      tma::prefetch_map(&p.e_map);
      tma::prefetch_map(&p.pj_map);
      tma::prefetch_map(&p.out_map);
    }
    const bool e_gather = p.e_index != nullptr, o_scatter = p.out_index != nullptr;
    // `keep`: the tile's out_index rows (scatter mode), held until the tile is stored
    auto load_e = [&](int tile, uint32_t slot_i, int4& keep) {
      const uint32_t row0 = (uint32_t)tile * EW_TM;
      const int rows_here = (int)min((int64_t)EW_TM, p.n_rows - (int64_t)row0);
      const uint32_t sl = slots + slot_i * EW_SLOT, fb = full0 + 8 * slot_i;
      int4 e4 = make_int4(0, 0, 0, 0);
      if (e_gather) e4 = ew_idx4(p.e_index, row0, rows_here, lane, 0);
      if (o_scatter) keep = (e_gather && p.e_index == p.out_index) ? e4 : ew_idx4(p.out_index, row0, rows_here, lane, 0);
      if (lane == 0) tma::mbar_expect_tx(fb, EW_SLOT);
      __syncwarp();
      if (e_gather) {
        ew_gather_tile(sl, &p.e_map, fb, e4);
      } else if (elect_one()) {
        tma::load_2d(sl, &p.e_map, fb, 0, (int)row0);
        tma::load_2d(sl + 16384, &p.e_map, fb, 32, (int)row0);
      }
      __syncwarp();
    };
    auto load_pj = [&](int tile, uint32_t slot_i) {
      const uint32_t row0 = (uint32_t)tile * EW_TM;
      const int rows_here = (int)min((int64_t)EW_TM, p.n_rows - (int64_t)row0);
      const uint32_t sl = slots + slot_i * EW_SLOT, fb = full0 + 8 * slot_i;
      const int4 s4 = ew_idx4(p.src, row0, rows_here, lane, 0);
      if (lane == 0) tma::mbar_expect_tx(fb, EW_SLOT);
      __syncwarp();
      ew_gather_tile(sl, &p.pj_map, fb, s4);
    };
    int4 keep_cur = make_int4(0, 0, 0, 0), keep_next = keep_cur;
    if (tile0 < p.n_tiles) {
      load_e(tile0, 0, keep_cur);
      load_pj(tile0, 1);
    }
    int t = 0;
    for (int tile = tile0; tile < p.n_tiles; tile += tstep, ++t) {
      const uint32_t se = (uint32_t)t & 1u, sp = se ^ 1u;
      const int k = 3 * t;  // item numbers of this tile: k (e_in), k + 1 (P_j), k + 2 (out); slot = item & 1
      const bool more = tile + tstep < p.n_tiles;
      if (more) {  // e_in(t + 1) = item k + 3 -> the slot P_j(t) = item k + 1 leaves after the first epilogue
        EW_PROF(16);
        ew_wait(empty0 + 8 * sp, (uint32_t)((k + 1) >> 1) & 1u);
        EW_PROF(17);
        load_e(tile + tstep, sp, keep_next);
        EW_PROF(18);
      }
      // out(t) = item k + 2, staged by the row owners in slot se
      ew_wait(out_ready, (uint32_t)t & 1u);
      EW_PROF(19);
      {
        const uint32_t row0 = (uint32_t)tile * EW_TM;
        const uint32_t sl = slots + se * EW_SLOT;
        const int rows_here = (int)min((int64_t)EW_TM, p.n_rows - (int64_t)row0);
        if (o_scatter) {
          if (rows_here == EW_TM) ew_scatter_tile(sl, &p.out_map, keep_cur);  // a partial tile is stored by its owners
        } else if (elect_one()) {
          tma::store_2d(&p.out_map, sl, 0, (int)row0);  // rows past the end of the table are clipped
          tma::store_2d(&p.out_map, sl + 16384, 32, (int)row0);
        }
        __syncwarp();
        tma::bulk_commit();
      }
      EW_PROF(20);
      if (more) {  // P_j(t + 1) = item k + 4 -> the slot out(t) leaves once summed and stored
        ew_wait(empty0 + 8 * se, (uint32_t)((k + 2) >> 1) & 1u);
        EW_PROF(21);
        tma::bulk_wait_read0();
        __syncwarp();
        EW_PROF(22);
        load_pj(tile + tstep, se);
        EW_PROF(23);
      }
      keep_cur = keep_next;
    }
    tma::bulk_wait_all0();
  } else if (warp < 4) {
    // ================================================================= MMA issue of context `ctx`
    const uint32_t idesc = make_idesc_tf32(EW_TM, 64);
    int n = 0;  // commits so far
    for (int tile = tile0; tile < p.n_tiles; tile += tstep) {
#pragma unroll 1
      for (int l = 0; l < 3; ++l, ++n) {
        EW_PROF(24);
        ew_wait(a_ready, (uint32_t)n & 1u);
        EW_PROF(25);
        tc_fence_after_sync();
        const uint64_t bd_hi = make_smem_desc_sw128(wbase + (uint32_t)l * 32768u);
        const uint64_t bd_lo = make_smem_desc_sw128(wbase + (uint32_t)l * 32768u + 16384u);
        if (elect_one()) {
          bool acc = false;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {  // small terms first: lo*hi, hi*lo, hi*hi
            const uint32_t a = tmc + ((pass == 0) ? EW_A_LO : EW_A_HI);
            const uint64_t bd = (pass == 1) ? bd_lo : bd_hi;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              mma_tf32_ts(tmc + EW_D, a + 8 * ks, bd + (uint64_t)((ks >> 2) * 512 + (ks & 3) * 2), idesc, acc);
              acc = true;
            }
          }
          mma_commit_addr(d_ready);
        }
        __syncwarp();
        EW_PROF(26);
      }
    }
  } else {
    // ================================================================= row owners of context `ctx`
    const int cw = (warp - 4) & 7;                 // warp inside the context
    const int r = 32 * (warp & 3) + lane, h = cw >> 2;  // TMEM lane quarter = warp % 4
    const int tt = 32 * cw + lane;
    const uint32_t tm_lane = tmc + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t rx = (uint32_t)(r & 7) << 4;
    const uint32_t own = (uint32_t)h * 16384u + (uint32_t)r * 128u;  // own row inside the own K tile of a slot
    const float* bias = reinterpret_cast<const float*>(smem_raw + EW_HEAD + 98304);
    const uint32_t pld4 = (uint32_t)p.pi_ld * 4u, ald4 = (uint32_t)p.aggr_ld * 4u;
    const int c4 = tt & 15, rg0 = (tt >> 4) * 8;   // segmented sum: 16-byte piece c4 of rows rg0 .. rg0 + 7
    int nd = 0;  // accumulator completions consumed so far
    int32_t dcur = 0;
    if (tile0 < p.n_tiles) {
      const uint32_t row = (uint32_t)tile0 * EW_TM + r;
      dcur = (int64_t)row < p.n_rows ? __ldg(p.pi_index + row) : 0;
    }
    int t = 0;
    for (int tile = tile0; tile < p.n_tiles; tile += tstep, ++t) {
      const uint32_t se = (uint32_t)t & 1u, sp = se ^ 1u;
      const uint32_t row0 = (uint32_t)tile * EW_TM;
      const int rows_here = (int)min((int64_t)EW_TM, p.n_rows - (int64_t)row0);
      const uint32_t sl_e = slots + se * EW_SLOT, sl_p = slots + sp * EW_SLOT;

      // ---------------- first Linear's A operand: own 32 columns of the edge-feature row
      EW_PROF(0);
      ew_wait(full0 + 8 * se, (uint32_t)t & 1u);
      EW_PROF(1);
      {
        float4 a[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = lds128(sl_e + own + (((uint32_t)q << 4) ^ rx));
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          float v[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            v[4 * q + 0] = a[4 * b + q].x; v[4 * q + 1] = a[4 * b + q].y;
            v[4 * q + 2] = a[4 * b + q].z; v[4 * q + 3] = a[4 * b + q].w;
          }
          if (RELU_E) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          split_store16(tm_lane + EW_A_HI + 32 * h + 16 * b, tm_lane + EW_A_LO + 32 * h + 16 * b, v);
        }
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        tma::mbar_arrive(a_ready);
        tma::mbar_arrive(empty0 + 8 * se);  // item e_in(t) consumed
      }
      EW_PROF(2);
      // gathered target rows (destination-sorted: neighbouring lanes repeat rows) and next tile's ids,
      // in flight under the first MMA chain
      float4 pre[8];
      {
        const float* rowp = row_ptr(p.pi + 32 * h, (uint32_t)dcur, pld4);
#pragma unroll
        for (int q = 0; q < 8; ++q) pre[q] = __ldg(reinterpret_cast<const float4*>(rowp) + q);
      }
      int32_t dnext = 0;
      if (tile + tstep < p.n_tiles) {
        const uint32_t row = (uint32_t)(tile + tstep) * EW_TM + r;
        dnext = (int64_t)row < p.n_rows ? __ldg(p.pi_index + row) : 0;
      }

      // ---------------- hidden layer 0: D + b0 + P_i[dst] + P_j[src] -> ReLU -> A operand
      EW_PROF(3);
      ew_wait(full0 + 8 * sp, (uint32_t)t & 1u);
      EW_PROF(4);
      ew_wait(d_ready, (uint32_t)nd & 1u);
      EW_PROF(5);
      ++nd;
      tc_fence_after_sync();
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int cb = 32 * h + 16 * b;
        uint32_t acc[16];
        tmem_ld16(tm_lane + EW_D + cb, acc);
        float4 x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = lds128(sl_p + own + (((uint32_t)(4 * b + q) << 4) ^ rx));
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
        const float4* b4 = reinterpret_cast<const float4*>(bias + cb);
        add16(v, b4[0], b4[1], b4[2], b4[3]);
        add16(v, pre[4 * b + 0], pre[4 * b + 1], pre[4 * b + 2], pre[4 * b + 3]);
        add16(v, x[0], x[1], x[2], x[3]);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
        split_store16(tm_lane + EW_A_HI + cb, tm_lane + EW_A_LO + cb, v);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        tma::mbar_arrive(a_ready);
        tma::mbar_arrive(empty0 + 8 * sp);  // item P_j(t) consumed
      }
      EW_PROF(6);

      // ---------------- hidden layer 1
      ew_wait(d_ready, (uint32_t)nd & 1u);
      EW_PROF(7);
      ++nd;
      tc_fence_after_sync();
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int cb = 32 * h + 16 * b;
        uint32_t acc[16];
        tmem_ld16(tm_lane + EW_D + cb, acc);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
        const float4* b4 = reinterpret_cast<const float4*>(bias + 64 + cb);
        add16(v, b4[0], b4[1], b4[2], b4[3]);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
        split_store16(tm_lane + EW_A_HI + cb, tm_lane + EW_A_LO + cb, v);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) tma::mbar_arrive(a_ready);
      EW_PROF(8);

      // segment ids of the rows this thread will sum (rg0 is a multiple of 8: 32-byte aligned loads)
      int sg[8];
      if (rows_here == EW_TM) {
        const int4 s0 = __ldg(reinterpret_cast<const int4*>(p.dst + row0 + rg0));
        const int4 s1 = __ldg(reinterpret_cast<const int4*>(p.dst + row0 + rg0) + 1);
        sg[0] = s0.x; sg[1] = s0.y; sg[2] = s0.z; sg[3] = s0.w; sg[4] = s1.x; sg[5] = s1.y; sg[6] = s1.z; sg[7] = s1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) sg[i] = rg0 + i < rows_here ? __ldg(p.dst + row0 + rg0 + i) : -1;
      }

      // ---------------- output Linear: D + b2 -> staged tile (slot se: the thread's own e_in piece is dead)
      EW_PROF(9);
      ew_wait(d_ready, (uint32_t)nd & 1u);
      EW_PROF(10);
      ++nd;
      tc_fence_after_sync();
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int cb = 32 * h + 16 * b;
        uint32_t acc[16];
        tmem_ld16(tm_lane + EW_D + cb, acc);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
        const float4* b4 = reinterpret_cast<const float4*>(bias + 128 + cb);
        add16(v, b4[0], b4[1], b4[2], b4[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          sts128(sl_e + own + (((uint32_t)(4 * b + q) << 4) ^ rx), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
      }
      if (p.out_index != nullptr && rows_here < EW_TM && r < rows_here) {  // partial last tile, scattered rows
        float* orow = const_cast<float*>(row_ptr(p.out + 32 * h, (uint32_t)__ldg(p.out_index + row0 + r), (uint32_t)p.out_ld * 4u));
#pragma unroll
        for (int q = 0; q < 8; ++q) reinterpret_cast<float4*>(orow)[q] = lds128(sl_e + own + (((uint32_t)q << 4) ^ rx));
      }
      fence_proxy_async_smem();  // the TMA store reads the tile through the async proxy
      __syncwarp();
      if (lane == 0) tma::mbar_arrive(out_ready);
      EW_PROF(11);

      // ---------------- in-tile segmented sum by destination: one vector reduction per run inside 8 rows
      ew_wait(out_ready, (uint32_t)t & 1u);
      EW_PROF(12);
      if (rg0 < rows_here) {
        float4 v[8];
        const uint32_t colb = sl_e + (uint32_t)(c4 >> 3) * 16384u + (uint32_t)rg0 * 128u;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = lds128(colb + (uint32_t)i * 128u + ((((uint32_t)c4 & 7u) ^ (uint32_t)i) << 4));
        int cur = sg[0];
        f32x2 s01 = pack2(0.f, 0.f), s23 = s01;
        auto flush = [&](int seg) {
          float4 sum;
          unpack2(s01, sum.x, sum.y);
          unpack2(s23, sum.z, sum.w);
          red_add_v4(row_ptr(p.aggr + 4 * c4, (uint32_t)seg, ald4), sum);
        };
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (sg[i] < 0) break;
          if (sg[i] != cur) {
            flush(cur);
            cur = sg[i];
            s01 = pack2(0.f, 0.f);
            s23 = s01;
          }
          s01 = add2(s01, pack2(v[i].x, v[i].y));
          s23 = add2(s23, pack2(v[i].z, v[i].w));
        }
        flush(cur);
      }
      __syncwarp();
      if (lane == 0) tma::mbar_arrive(empty0 + 8 * se);  // item out(t) consumed by the owners
      EW_PROF(13);
      if (PROF && prof_on) g_ew_prof[15] += 1;
      dcur = dnext;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0u, 512);
}

// ------------------------------------------------------------------------------ host side
// Does this descriptor have the shape of the wide IN edge kernel?  (Checked by fused_mlp_tc before the
// generic tiles.)  Returns the positions of the three source blocks through the out arguments.
static bool ew_match(const gtb_mlp_desc_t& d, int* s_e, int* s_pi, int* s_pj) {
  if (d.n_layers != 3 || d.n_srcs != 3) return false;
  for (int l = 0; l <= 3; ++l)
    if (d.dims[l] != 64) return false;
  if (d.aggr == nullptr || d.seg_id == nullptr || d.out == nullptr) return false;
  if (d.res != nullptr || d.res_b != 1.f || d.row_scale || d.out_scale || d.gate || d.final_act != GTB_ACT_NONE) return false;
  *s_e = *s_pi = *s_pj = -1;
  for (int s = 0; s < 3; ++s) {
    const gtb_src_t& b = d.srcs[s];
    if (b.width != 64 || (b.ld & 3) || (reinterpret_cast<uintptr_t>(b.ptr) & 15)) return false;
    if (b.flags & GTB_SRC_PROJECTED) {
      if (b.relu || b.index == nullptr) return false;
      if ((b.index == d.seg_id || (b.flags & GTB_SRC_SORTED)) && *s_pi < 0) *s_pi = s;  // gathered by the sorted destination ids
      else if (*s_pj < 0) *s_pj = s;
      else return false;
    } else {
      if (*s_e >= 0) return false;
      *s_e = s;
    }
  }
  if (*s_e < 0 || *s_pi < 0 || *s_pj < 0) return false;
  if ((d.aggr_ld & 3) || (reinterpret_cast<uintptr_t>(d.aggr) & 15)) return false;
  if ((d.out_ld & 3) || (reinterpret_cast<uintptr_t>(d.out) & 15)) return false;
  // index arrays are read as int4 per tile
  if ((reinterpret_cast<uintptr_t>(d.seg_id) & 15) || (reinterpret_cast<uintptr_t>(d.srcs[*s_pj].index) & 15)) return false;
  if (d.srcs[*s_e].index && (reinterpret_cast<uintptr_t>(d.srcs[*s_e].index) & 15)) return false;
  if (d.out_index && (reinterpret_cast<uintptr_t>(d.out_index) & 15)) return false;
  return true;
}

// n_table_rows: rows of the gathered tables are not part of the C descriptor; the TMA map only needs an
// upper bound for its bounds check, the indices are the plan's (validated when the plan is built).
int in_edge_ws(const gtb_mlp_desc_t& d, cudaStream_t st, bool* handled) {
  *handled = false;
  static const bool disabled = getenv("GTB_NO_EDGE_WS") != nullptr;
  int s_e, s_pi, s_pj;
  if (disabled || !ew_match(d, &s_e, &s_pi, &s_pj)) return GTB_OK;
  if (tma::encode_fn() == nullptr) return GTB_OK;
  EwParams p;
  memset(&p, 0, sizeof(p));
  const gtb_src_t &se = d.srcs[s_e], &spi = d.srcs[s_pi], &spj = d.srcs[s_pj];
  const uint64_t big = 1ull << 31;  // row bound of gathered tables (indices come from a validated plan)
  const bool e_gather = se.index != nullptr, o_scatter = d.out_index != nullptr;
  const uint64_t e_rows = e_gather ? big : (uint64_t)d.n_rows, o_rows = o_scatter ? big : (uint64_t)d.n_rows;
  if (!tma::make_map_2d(&p.e_map, se.ptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, e_rows, 64, (uint64_t)se.ld, 32, e_gather ? 1 : 128) ||
      !tma::make_map_2d(&p.pj_map, spj.ptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, big, 64, (uint64_t)spj.ld, 32, 1) ||
      !tma::make_map_2d(&p.out_map, d.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, o_rows, 64, (uint64_t)d.out_ld, 32, o_scatter ? 1 : 128))
    return GTB_OK;  // the driver refused a map: the generic tiles take the launch
  p.pi = spi.ptr;
  p.pi_ld = spi.ld;
  p.dst = d.seg_id;
  p.pi_index = spi.index;
  p.src = spj.index;
  p.e_index = se.index;
  p.out_index = d.out_index;
  p.aggr = d.aggr;
  p.aggr_ld = d.aggr_ld;
  p.packed = static_cast<const unsigned char*>(d.packed);
  p.w_bytes = 99072;
  p.n_rows = d.n_rows;
  p.n_tiles = (int32_t)((d.n_rows + EW_TM - 1) / EW_TM);
  p.out = d.out;
  p.out_ld = d.out_ld;
  *handled = true;
  if (d.n_rows == 0) return GTB_OK;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(in_edge_ws_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EW_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(in_edge_ws_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EW_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(in_edge_ws_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EW_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(in_edge_ws_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EW_SMEM);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(in_edge_ws)");
    configured[dev] = true;
  }
  const int pairs = (p.n_tiles + 1) / 2;
  const int grid = pairs < kNumSMs ? pairs : kNumSMs;
  if (g_ew_prof_enabled) {
    if (se.relu) in_edge_ws_kernel<true, true><<<grid, EW_THREADS, EW_SMEM, st>>>(p);
    else         in_edge_ws_kernel<false, true><<<grid, EW_THREADS, EW_SMEM, st>>>(p);
  } else {
    if (se.relu) in_edge_ws_kernel<true, false><<<grid, EW_THREADS, EW_SMEM, st>>>(p);
    else         in_edge_ws_kernel<false, false><<<grid, EW_THREADS, EW_SMEM, st>>>(p);
  }
  GTB_CHECK_LAUNCH("in_edge_ws_kernel");
  return GTB_OK;
}

int ew_profile(int enable, long long* out32) {
  if (out32 == nullptr) {
    g_ew_prof_enabled = enable;
    long long zero[32] = {0};
    return check_cuda(cudaMemcpyToSymbol(g_ew_prof, zero, sizeof(zero)), "cudaMemcpyToSymbol(g_ew_prof)");
  }
  return check_cuda(cudaMemcpyFromSymbol(out32, g_ew_prof, 32 * sizeof(long long)), "cudaMemcpyFromSymbol(g_ew_prof)");
}

int ew_fault_flag(int* out) { return check_cuda(cudaMemcpyFromSymbol(out, g_ew_fault, sizeof(int)), "cudaMemcpyFromSymbol(g_ew_fault)"); }

}  // namespace gtb
