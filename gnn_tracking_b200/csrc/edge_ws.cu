// Warp-specialised fused Interaction-Network edge kernel (sm_100a), the 64 / 64 / 64 "wide" shape:
//
//   e_out[o(r)] = W2 relu(W1 relu(W0e act(e_in[i(r)]) + P_i[dst(r)] + P_j[src(r)] + b0) + b1) + b2
//   aggr[dst(r)] += e_out row          (rows r walk the plan's destination-sorted edge list)
//
// i.e. reference models/interaction_network.py:75-89 (message) + the SumAggregation of :22,36 with the
// node-side products P_i = act(x) W0[:, :Dn]^T, P_j = act(x) W0[:, Dn:2Dn]^T taken per node (see
// GTB_SRC_PROJECTED in include/gtb200.h).  Same arithmetic as the generic tiles of mlp_tc.cu (3xTF32,
// A operand in TMEM, fp32 accumulation), different machine mapping:
//
//   * one persistent CTA per SM, two tiles ("contexts" A / B) in flight, 20 warps with fixed roles:
//       warps 0-15  : 512 row owners (thread = row x 16-column quarter).  ALL of them work on one
//                     stage of one context at a time and ping-pong between the contexts: while the
//                     tensor core runs a Linear of context A they run a stage of context B --
//                        E0(A) E0(B) E1(A) E1(B) E2(A) E2(B) AG(A) RF(A) C0(A') AG(B) RF(B) C0(B')
//                     C0: tf32 hi / lo split of the edge-feature tile into TMEM; E0 / E1: accumulator
//                     + bias (+ P_i[dst] + P_j[src]) -> ReLU -> next A operand; E2: output tile into
//                     shared memory; AG: store of the tile, in-tile segmented sum by destination, and
//                     the gather of the NEXT tile's P_j rows into the rows just vacated
//       warp 16 / 17: TMA producer of context A / B: edge-feature tile (2-D tile load when the features
//                     are kept in destination order, tile::gather4 through `perm` otherwise) and the
//                     tile's dst / src ids (bulk copies)
//       warp 18 / 19: tcgen05.mma issue of context A / B (one elected lane, 24 MMAs per Linear)
//   * every hand-over is an mbarrier (no CTA or named barrier inside the tile loop):
//       full_e   producer -> owners  (edge-feature tile + ids landed: transaction bytes)
//       full_p   owners   -> owners  (P_j rows landed: 16 warps x their own 8 rows, transaction bytes)
//       pj_free  owners   -> producer (P_j consumed: its slot takes the next edge-feature tile)
//       a_ready  owners   -> MMA warp (A operand of the next Linear is in TMEM)
//       d_ready  MMA warp -> owners   (tcgen05.commit: accumulator complete, A operand free)
//       out_ready owners  -> owners   (output tile staged: rows regroup from row owners to 8-row bands)
//   * TMA: gathers are tile::gather4 (four row coordinates per instruction) issued from single-lane
//     branches with broadcast coordinates; stores are 8-row tile stores / tile::scatter4, issued by the
//     warp that sums the same 8 rows, so a band of a slot is refilled by the warp that drained it and the
//     slot ring needs no producer round trip: per context two 32 KB slots carry e_in(t), P_j(t), out(t).
//   * shared memory: weights 96 KB (3 x hi / lo x 16 KB, the image gtb_mlp_pack writes) | 4 slots |
//     biases, ids, barriers = 227 KB.  TMEM: 192 columns per context (A hi | A lo | D).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace gtb {

using namespace tc;

constexpr int EW_TM = 128;
constexpr int EW_SLOT = 32768;        // [2 K tiles][128 rows][32 fp32], 128-byte swizzle (TMA image == UMMA image)
constexpr int EW_W = 98304;           // packed weights without the bias rows
constexpr int EW_SLOTS = EW_W;        // 4 slots: context c, slot s at EW_SLOTS + (2 c + s) * EW_SLOT
constexpr int EW_BIAS = EW_SLOTS + 4 * EW_SLOT;  // 3 x 64 floats
constexpr int EW_IDS = EW_BIAS + 768;  // per context: dst ids [128] | src ids [128]
constexpr int EW_BARS = EW_IDS + 2048;  // per context 64 bytes: full_e | full_p | pj_free | a_ready | d_ready | out_ready
constexpr int EW_TMEM_SLOT = EW_BARS + 128;
constexpr int EW_SMEM = 232448;       // the 227 KB opt-in maximum
static_assert(EW_TMEM_SLOT + 4 <= EW_SMEM, "shared-memory layout");
constexpr int EW_THREADS = 640;
constexpr uint32_t EW_A_HI = 0, EW_A_LO = 64, EW_D = 128, EW_CTX = 192;

__device__ int g_ew_fault = 0;  // 1: barrier timeout, 2: TMEM base != 0, 3: shared memory misaligned
__device__ long long g_ew_prof[32];
static int g_ew_prof_enabled = 0;

// per-stage clock accumulation by lane 0 of one warp per role in CTA 0 (tests/cuda/tc_diag.py)
#define EW_PROF(id)                                   \
  do {                                                \
    if (PROF && prof_on) {                            \
      const long long now_ = clock64();               \
      g_ew_prof[id] += now_ - prof_t;                 \
      prof_t = now_;                                  \
    }                                                 \
  } while (0)

struct EwParams {
  CUtensorMap e_map;     // e_in  [E, 64] fp32: box 32 x 128 (tile mode) or 32 x 1 (gather mode)
  CUtensorMap pj_map;    // P_j   [N, 64] fp32: box 32 x 1
  CUtensorMap out_map;   // e_out [E, 64] fp32: box 32 x 8 (tile mode) or 32 x 1 (scatter mode)
  const void* pi;        // P_i [N, pi_ld], rows addressed by dst (fp32, or bf16 in the bf16 variant)
  const int32_t* dst;    // dst_sorted [E]: P_i row and segment id of every edge
  const int32_t* src;    // src_sorted [E]
  const int32_t* e_index;    // perm or nullptr (tile mode)
  const int32_t* out_index;  // perm or nullptr (tile mode)
  float* aggr;
  void* out;             // e_out base (scatter mode: a partial last tile is stored with plain stores)
  const unsigned char* packed;
  int64_t n_rows;
  int32_t n_tiles, pi_ld, aggr_ld, out_ld;
  float* h0;             // SAVE variant: post-ReLU outputs of the first / second Linear, [n_rows, h_ld] in launch-row order
  float* h1;
  int32_t h_ld;
  int32_t debug_bar;     // GTB_EW_DEBUG_BAR=1: a named barrier beside the out_ready mbarrier (compute-sanitizer racecheck
                         // does not model mbarrier hand-overs; with the barrier in place it must report no hazard)
};

__device__ __noinline__ void ew_timeout();
__device__ __forceinline__ void ew_wait(uint32_t bar, uint32_t parity) {
  // First try outside the loop: mbarrier.try_wait blocks in hardware for a while, and a load issued in front of
  // the wait (P_i[dst] in E0) has its scoreboard wait parked at the next basic-block boundary -- with the first
  // try peeled that boundary lies BEHIND one blocking try, so the load's latency runs under the barrier wait.
  {
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
  // NOT unrolled: ptxas otherwise replicates the try_wait 64 times per wait site (100 KB of SASS: every
  // stage change then misses the instruction cache)
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
  ew_timeout();
}

// two barriers at once: both (blocking) first tries in straight-line code before the first branch, so that loads issued
// in front of the call stay in flight under BOTH waits
__device__ __forceinline__ void ew_wait2(uint32_t bar_a, uint32_t par_a, uint32_t bar_b, uint32_t par_b) {
  uint32_t ok_a, ok_b;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok_a)
      : "r"(bar_a), "r"(par_a)
      : "memory");
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok_b)
      : "r"(bar_b), "r"(par_b)
      : "memory");
  if (ok_a & ok_b) return;
  if (!ok_a) ew_wait(bar_a, par_a);
  if (!ok_b) ew_wait(bar_b, par_b);
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// row indices 4 lane .. 4 lane + 3 of a tile (`fill` past the end of the list)
__device__ __forceinline__ int4 ew_idx4(const int32_t* idx, uint32_t row0, int rows_here, int lane, int fill) {
  if (rows_here == EW_TM) return __ldg(reinterpret_cast<const int4*>(idx + row0) + lane);
  int4 v;
  v.x = 4 * lane + 0 < rows_here ? __ldg(idx + row0 + 4 * lane + 0) : fill;
  v.y = 4 * lane + 1 < rows_here ? __ldg(idx + row0 + 4 * lane + 1) : fill;
  v.z = 4 * lane + 2 < rows_here ? __ldg(idx + row0 + 4 * lane + 2) : fill;
  v.w = 4 * lane + 3 < rows_here ? __ldg(idx + row0 + 4 * lane + 3) : fill;
  return v;
}

// Barrier timeouts leave through one out-of-line exit (the wait loop itself stays three instructions).
__device__ __noinline__ void ew_timeout() {
  atomicExch(&g_ew_fault, 1);
  __trap();  // a protocol error must fail loudly, never hang the GPU
}

// state of the 512 row owners.  The context c is a RUN-TIME value: one copy of every stage serves both
// contexts (the loop body stays inside the instruction cache); the only per-context registers are the
// four segment ids a lane sums in AG, selected with c.
template <bool BF, bool RELU_E, bool PROF, bool SAVE>
struct EwOwner {
  static constexpr uint32_t ES = BF ? 2u : 4u;            // bytes per element of the tables
  static constexpr int KT = BF ? 64 : 32;                 // columns of one 128-byte K tile
  static constexpr uint32_t TA = 0, TD = BF ? 64u : 128u;  // TMEM columns: A operand (bf16: 64 columns; fp32: hi | lo), accumulator

  const EwParams& p;
  uint32_t sm0;
  int w, lane, r, qd;       // warp 0..15, row 32 (w & 3) + lane, column quarter w >> 2
  uint32_t rx, own;         // swizzle term of the own row, own row inside the own K tile of a slot
  int tile00, nA, nB;       // first tile of context 0, number of tiles of each context (stride 2 * gridDim.x)
  int32_t sgA[4], sgB[4];   // segment ids of the 4 rows this lane sums
  bool prof_on;
  long long prof_t;

  __device__ __forceinline__ int n_of(int c) const { return c ? nB : nA; }
  __device__ __forceinline__ int tile_of(int c, int t) const { return tile00 + (2 * t + c) * (int)gridDim.x; }
  __device__ __forceinline__ uint32_t bar(int c, int which) const { return sm0 + EW_BARS + 64 * c + 8 * which; }
  __device__ __forceinline__ uint32_t slot(int c, uint32_t s) const { return sm0 + EW_SLOTS + (2 * c + s) * EW_SLOT; }
  __device__ __forceinline__ uint32_t tm_lane(int c) const { return (uint32_t)c * EW_CTX + ((uint32_t)((w & 3) * 32) << 16); }
  __device__ __forceinline__ uint32_t chunk(int q) const { return (((uint32_t)(4 * (qd & 1) + q)) << 4) ^ rx; }
  __device__ __forceinline__ uint32_t ids(int c) const { return sm0 + EW_IDS + 1024 * c; }
  __device__ __forceinline__ void arrive(uint32_t b) {
    __syncwarp();
    if (lane == 0) tma::mbar_arrive(b);
  }
  __device__ __forceinline__ void a_done(int c) {  // the A operand written by this warp is complete
    tmem_st_wait();
    tc_fence_before_sync();
    arrive(bar(c, 3));
  }
  __device__ __forceinline__ static float4 u2f(const uint4& u) {
    return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
  }
  // own quarter (64 bytes) of row d of P_i
  __device__ __forceinline__ const void* pi_row(int32_t d) const {
    return reinterpret_cast<const char*>(p.pi) + (uint64_t)(uint32_t)d * ((uint32_t)p.pi_ld * ES) + 64u * (uint32_t)qd;
  }
  // ---- bf16 variant: own 32 columns
  __device__ __forceinline__ void acc_load32(int c, int t, int l, float (&v)[32], bool waited = false) {
    if (!waited) ew_wait(bar(c, 4), (uint32_t)(t + l) & 1u);
    tc_fence_after_sync();
    uint32_t a0[16], a1[16];
    tmem_ld16(tm_lane(c) + TD + 32 * qd, a0);
    tmem_ld16(tm_lane(c) + TD + 32 * qd + 16, a1);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v[j] = __uint_as_float(a0[j]);
      v[16 + j] = __uint_as_float(a1[j]);
    }
  }
  __device__ __forceinline__ void bias_add32(float (&v)[32], int layer) {
    const uint32_t ba = sm0 + EW_BIAS + 256 * layer + 64 * qd;
    bf2_add16(v, 0, lds128u(ba), lds128u(ba + 16));
    bf2_add16(v, 16, lds128u(ba + 32), lds128u(ba + 48));
  }
  // Linear output rounded to bf16 (what autocast's Linear returns), ReLU, into the next A operand
  __device__ __forceinline__ void act_store32(int c, const float (&v)[32]) {
    uint32_t a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = bf2_relu(bf2_pack_rn(v[2 * j], v[2 * j + 1]));
    tmem_st16(tm_lane(c) + TA + 16 * qd, a);
  }
  __device__ __forceinline__ void bias_add(float (&v)[16], int layer) {
    const uint32_t ba = sm0 + EW_BIAS + 256 * layer + 64 * qd;
    add16(v, lds128(ba), lds128(ba + 16), lds128(ba + 32), lds128(ba + 48));
  }
  // accumulator of the own 16 columns, once the l-th Linear of tile iteration t has completed
  __device__ __forceinline__ void acc_load(int c, int t, int l, float (&v)[16], bool waited = false) {
    if (!waited) ew_wait(bar(c, 4), (uint32_t)(t + l) & 1u);  // completion number 3 t + l
    tc_fence_after_sync();
    uint32_t acc[16];
    tmem_ld16(tm_lane(c) + EW_D + 16 * qd, acc);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
  }

  // SAVE: the own 16 post-ReLU columns of row (tile, r) into a [n_rows, h_ld] table (the backward pass reads them
  // instead of recomputing the hidden activations)
  __device__ __forceinline__ void save16(float* table, int c, int t, const float (&v)[16]) const {
    const int64_t row = (int64_t)tile_of(c, t) * EW_TM + r;
    if (row < p.n_rows) {
      float4* q = reinterpret_cast<float4*>(table + row * (int64_t)p.h_ld + 16 * qd);
#pragma unroll
      for (int i = 0; i < 4; ++i) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
  }

  // ---- C0: own 16 columns of the edge-feature row -> tf32 hi / lo -> TMEM
  __device__ __forceinline__ void c0(int c, int t) {
    const int tile = tile_of(c, t);
    const uint32_t row0 = (uint32_t)tile * EW_TM;
    const int rows_here = (int)min((int64_t)EW_TM, p.n_rows - (int64_t)row0);
    const uint32_t sl = slot(c, (uint32_t)t & 1u);
    EW_PROF(0);
    ew_wait(bar(c, 0), (uint32_t)t & 1u);
    EW_PROF(1);
    const char* pi_own;  // own 64 bytes of the target row of P_i
    {
      int32_t d = 0;
      if (rows_here == EW_TM) d = lds_i32(ids(c) + 4 * r);
      else if (r < rows_here) d = __ldg(p.dst + row0 + r);
      pi_own = reinterpret_cast<const char*>(pi_row(d));
    }
    if constexpr (BF) {  // 32 bf16 columns = 16 words, straight into the A operand
      uint32_t a[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 x = lds128u(sl + own + chunk(q));
        a[4 * q + 0] = x.x; a[4 * q + 1] = x.y; a[4 * q + 2] = x.z; a[4 * q + 3] = x.w;
      }
      if (RELU_E) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = bf2_relu(a[j]);
      }
      tmem_st16(tm_lane(c) + TA + 16 * qd, a);
    } else {
      float v[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 a = lds128(sl + own + chunk(q));
        v[4 * q + 0] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
      }
      if (RELU_E) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      split_store16(tm_lane(c) + EW_A_HI + 16 * qd, tm_lane(c) + EW_A_LO + 16 * qd, v);
    }
    a_done(c);
    prefetch_l1(pi_own);  // read in E0, a Linear from now
    EW_PROF(2);
  }

  // ---- E0: D + b0 + P_i[dst] + P_j[src] -> ReLU -> A operand of the second Linear
  __device__ __forceinline__ void e0(int c, int t) {
    const int tile = tile_of(c, t);
    const uint32_t row0 = (uint32_t)tile * EW_TM;
    const int rows_here = (int)min((int64_t)EW_TM, p.n_rows - (int64_t)row0);
    const uint32_t sl = slot(c, ((uint32_t)t & 1u) ^ 1u);
    {  // segment ids of the rows this lane sums in AG: rows 8 w + 4 (lane >> 4) + i (the ids leave with P_j).
       // In FRONT of the P_i loads and without a branch on the context: a branch behind a load waits for it.
      const int rr0 = 8 * w + 4 * (lane >> 4);
      int4 s;
      if (rows_here == EW_TM) {
        asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(s.x), "=r"(s.y), "=r"(s.z), "=r"(s.w) : "r"(ids(c) + 4 * rr0) : "memory");
      } else {
        s.x = rr0 + 0 < rows_here ? __ldg(p.dst + row0 + rr0 + 0) : -1;
        s.y = rr0 + 1 < rows_here ? __ldg(p.dst + row0 + rr0 + 1) : -1;
        s.z = rr0 + 2 < rows_here ? __ldg(p.dst + row0 + rr0 + 2) : -1;
        s.w = rr0 + 3 < rows_here ? __ldg(p.dst + row0 + rr0 + 3) : -1;
      }
      const bool cb = c != 0;
      sgA[0] = cb ? sgA[0] : s.x; sgA[1] = cb ? sgA[1] : s.y; sgA[2] = cb ? sgA[2] : s.z; sgA[3] = cb ? sgA[3] : s.w;
      sgB[0] = cb ? s.x : sgB[0]; sgB[1] = cb ? s.y : sgB[1]; sgB[2] = cb ? s.z : sgB[2]; sgB[3] = cb ? s.w : sgB[3];
    }
    uint4 pre[4];  // own 64 bytes of P_i[dst]: destination-sorted rows, neighbouring lanes repeat lines
    {
      int32_t d = 0;
      if (rows_here == EW_TM) d = lds_i32(ids(c) + 4 * r);
      else if (r < rows_here) d = __ldg(p.dst + row0 + r);
      const uint4* rowp = reinterpret_cast<const uint4*>(pi_row(d));
#pragma unroll
      for (int q = 0; q < 4; ++q) pre[q] = __ldg(rowp + q);
    }
    EW_PROF(3);
    ew_wait2(bar(c, 1), (uint32_t)t & 1u, bar(c, 4), (uint32_t)t & 1u);  // P_j rows landed; first Linear complete
    EW_PROF(4);
    if constexpr (BF) {
      float v[32];
      acc_load32(c, t, 0, v, true);
      EW_PROF(5);
      uint4 x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = lds128u(sl + own + chunk(q));
      bias_add32(v, 0);
      bf2_add16(v, 0, pre[0], pre[1]);
      bf2_add16(v, 16, pre[2], pre[3]);
      bf2_add16(v, 0, x[0], x[1]);
      bf2_add16(v, 16, x[2], x[3]);
      act_store32(c, v);
    } else {
      float v[16];
      acc_load(c, t, 0, v, true);
      EW_PROF(5);
      float4 x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = lds128(sl + own + chunk(q));
      bias_add(v, 0);
      add16(v, u2f(pre[0]), u2f(pre[1]), u2f(pre[2]), u2f(pre[3]));
      add16(v, x[0], x[1], x[2], x[3]);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
      split_store16(tm_lane(c) + EW_A_HI + 16 * qd, tm_lane(c) + EW_A_LO + 16 * qd, v);
      if constexpr (SAVE) save16(p.h0, c, t, v);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) {
      tma::mbar_arrive(bar(c, 3));
      tma::mbar_arrive(bar(c, 2));  // P_j(t) consumed: the producer may load e_in(t + 1) over it
    }
    EW_PROF(6);
  }

  // ---- E1: D + b1 -> ReLU -> A operand of the third Linear
  __device__ __forceinline__ void e1(int c, int t) {
    EW_PROF(7);
    if constexpr (BF) {
      float v[32];
      acc_load32(c, t, 1, v);
      EW_PROF(8);
      bias_add32(v, 1);
      act_store32(c, v);
    } else {
      float v[16];
      acc_load(c, t, 1, v);
      EW_PROF(8);
      bias_add(v, 1);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
      split_store16(tm_lane(c) + EW_A_HI + 16 * qd, tm_lane(c) + EW_A_LO + 16 * qd, v);
      if constexpr (SAVE) save16(p.h1, c, t, v);
    }
    a_done(c);
    EW_PROF(9);
  }

  // ---- E2: D + b2 -> output tile in shared memory (the slot of e_in(t): the own piece is dead)
  __device__ __forceinline__ void e2(int c, int t) {
    const uint32_t sl = slot(c, (uint32_t)t & 1u);
    EW_PROF(10);
    if constexpr (BF) {
      float v[32];
      acc_load32(c, t, 2, v);
      EW_PROF(11);
      bias_add32(v, 2);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        sts128u(sl + own + chunk(q), make_uint4(bf2_pack_rn(v[8 * q + 0], v[8 * q + 1]), bf2_pack_rn(v[8 * q + 2], v[8 * q + 3]),
                                                bf2_pack_rn(v[8 * q + 4], v[8 * q + 5]), bf2_pack_rn(v[8 * q + 6], v[8 * q + 7])));
    } else {
      float v[16];
      acc_load(c, t, 2, v);
      EW_PROF(11);
      bias_add(v, 2);
#pragma unroll
      for (int q = 0; q < 4; ++q) sts128(sl + own + chunk(q), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    }
    tc_fence_before_sync();
    fence_proxy_async_smem();  // the TMA store reads the tile through the async proxy
    arrive(bar(c, 5));
    EW_PROF(12);
  }

  // gather of the P_j rows 8 w .. 8 w + 7 of tile iteration `t` into `sl` (this warp's band of the slot)
  __device__ __forceinline__ void load_pj(int c, int t, uint32_t sl) {
    const int tile = tile_of(c, t);
    const uint32_t row0 = (uint32_t)tile * EW_TM;
    const int rows_here = (int)min((int64_t)EW_TM, p.n_rows - (int64_t)row0);
    ew_wait(bar(c, 0), (uint32_t)t & 1u);  // the ids of tile t travel with its edge-feature tile
    const uint32_t fb = bar(c, 1);
    const uint32_t d0 = sl + (uint32_t)(8 * w) * 128u;
    if (rows_here == EW_TM) {
      // one lane: eight ids from shared memory -> uniform registers -> four gather4 (no shuffles, no re-convergence)
      if (elect_one()) {
        int4 s0, s1;
        const uint32_t ia = ids(c) + 512 + 32 * w;
        asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(s0.x), "=r"(s0.y), "=r"(s0.z), "=r"(s0.w) : "r"(ia) : "memory");
        asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(s1.x), "=r"(s1.y), "=r"(s1.z), "=r"(s1.w) : "r"(ia + 16) : "memory");
        tma::mbar_expect_tx(fb, 8 * 256);
        tma::gather4(d0, &p.pj_map, fb, 0, s0.x, s0.y, s0.z, s0.w);
        tma::gather4(d0 + 16384u, &p.pj_map, fb, KT, s0.x, s0.y, s0.z, s0.w);
        tma::gather4(d0 + 512u, &p.pj_map, fb, 0, s1.x, s1.y, s1.z, s1.w);
        tma::gather4(d0 + 16384u + 512u, &p.pj_map, fb, KT, s1.x, s1.y, s1.z, s1.w);
      }
      __syncwarp();
      return;
    }
    int32_t s = 0;  // partial last tile: guarded loads, rows past the end gather row 0
    if (lane < 8 && 8 * w + lane < rows_here) s = __ldg(p.src + row0 + 8 * w + lane);
    if (lane == 0) tma::mbar_expect_tx(fb, 8 * 256);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int a = __shfl_sync(0xffffffffu, s, 4 * j + 0), b = __shfl_sync(0xffffffffu, s, 4 * j + 1);
      const int cc = __shfl_sync(0xffffffffu, s, 4 * j + 2), d = __shfl_sync(0xffffffffu, s, 4 * j + 3);
      if (elect_one()) {
        tma::gather4(d0 + (uint32_t)(4 * j) * 128u, &p.pj_map, fb, 0, a, b, cc, d);
        tma::gather4(d0 + 16384u + (uint32_t)(4 * j) * 128u, &p.pj_map, fb, KT, a, b, cc, d);
      }
      __syncwarp();
    }
  }

  // ---- AG: rows regroup into 8-row bands (warp w: rows 8 w .. 8 w + 7, all 64 columns): store of the band,
  // segmented sum by destination, then the band takes the next tile's P_j rows
  __device__ __forceinline__ void ag(int c, int t) {
    const int tile = tile_of(c, t);
    const uint32_t row0 = (uint32_t)tile * EW_TM;
    const int rows_here = (int)min((int64_t)EW_TM, p.n_rows - (int64_t)row0);
    const uint32_t sl = slot(c, (uint32_t)t & 1u);
    EW_PROF(13);
    ew_wait(bar(c, 5), (uint32_t)t & 1u);
    if (p.debug_bar) asm volatile("bar.sync 1, 512;" ::: "memory");
    EW_PROF(14);
    if (8 * w < rows_here) {
      if (p.out_index == nullptr) {
        if (elect_one()) {  // rows past the end of the table are clipped
          tma::store_2d(&p.out_map, sl + (uint32_t)(8 * w) * 128u, 0, (int)row0 + 8 * w);
          tma::store_2d(&p.out_map, sl + 16384u + (uint32_t)(8 * w) * 128u, KT, (int)row0 + 8 * w);
        }
        __syncwarp();
      } else if (rows_here == EW_TM) {
        const int32_t o = lane < 8 ? __ldg(p.out_index + row0 + 8 * w + lane) : 0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int a = __shfl_sync(0xffffffffu, o, 4 * j + 0), b = __shfl_sync(0xffffffffu, o, 4 * j + 1);
          const int cc = __shfl_sync(0xffffffffu, o, 4 * j + 2), d = __shfl_sync(0xffffffffu, o, 4 * j + 3);
          if (elect_one()) {
            tma::scatter4(&p.out_map, sl + (uint32_t)(8 * w + 4 * j) * 128u, 0, a, b, cc, d);
            tma::scatter4(&p.out_map, sl + 16384u + (uint32_t)(8 * w + 4 * j) * 128u, KT, a, b, cc, d);
          }
          __syncwarp();
        }
      } else {  // partial last tile, scattered rows: plain stores
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = 8 * w + 2 * i + (lane >> 4), cc = lane & 15;
          if (row < rows_here) {
            const uint4 v = lds128u(sl + (uint32_t)(cc >> 3) * 16384u + (uint32_t)row * 128u + ((((uint32_t)cc & 7u) ^ ((uint32_t)row & 7u)) << 4));
            char* orow = reinterpret_cast<char*>(p.out) + (uint64_t)(uint32_t)__ldg(p.out_index + row0 + row) * ((uint32_t)p.out_ld * ES);
            reinterpret_cast<uint4*>(orow)[cc] = v;
          }
        }
      }
      tma::bulk_commit();
      // one vector reduction (red.global.add.v4.f32) per run of equal destinations inside 4 rows
      const int c4 = lane & 15, rr0 = 8 * w + 4 * (lane >> 4);
      uint4 v[4];
      const uint32_t colb = sl + (uint32_t)(c4 >> 3) * 16384u + (uint32_t)rr0 * 128u;
#pragma unroll
      for (int i = 0; i < 4; ++i)  // (rr0 + i) & 7 == 4 (lane >> 4) + i
        v[i] = lds128u(colb + (uint32_t)i * 128u + ((((uint32_t)c4 & 7u) ^ (uint32_t)(4 * (lane >> 4) + i)) << 4));
      int32_t sg[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) sg[i] = c ? sgB[i] : sgA[i];
      if (sg[0] >= 0) {
        const uint32_t ald4 = (uint32_t)p.aggr_ld * 4u;
        int cur = sg[0];
        if constexpr (BF) {  // the 16-byte piece holds 8 bf16 columns: fp32 sums, two vector reductions per run
          float s8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) s8[j] = 0.f;
          auto flush = [&](int seg) {
            const float* a = row_ptr(p.aggr + 8 * c4, (uint32_t)seg, ald4);
            red_add_v4(a, make_float4(s8[0], s8[1], s8[2], s8[3]));
            red_add_v4(a + 4, make_float4(s8[4], s8[5], s8[6], s8[7]));
          };
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (sg[i] < 0) break;
            if (sg[i] != cur) {
              flush(cur);
              cur = sg[i];
#pragma unroll
              for (int j = 0; j < 8; ++j) s8[j] = 0.f;
            }
            const uint32_t wv[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float lo, hi;
              bf2_unpack(wv[j], lo, hi);
              s8[2 * j] += lo;
              s8[2 * j + 1] += hi;
            }
          }
          flush(cur);
        } else {
          f32x2 s01 = pack2(0.f, 0.f), s23 = s01;
          auto flush = [&](int seg) {
            float4 sum;
            unpack2(s01, sum.x, sum.y);
            unpack2(s23, sum.z, sum.w);
            red_add_v4(row_ptr(p.aggr + 4 * c4, (uint32_t)seg, ald4), sum);
          };
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (sg[i] < 0) break;
            if (sg[i] != cur) {
              flush(cur);
              cur = sg[i];
              s01 = pack2(0.f, 0.f);
              s23 = s01;
            }
            s01 = add2(s01, pack2(__uint_as_float(v[i].x), __uint_as_float(v[i].y)));
            s23 = add2(s23, pack2(__uint_as_float(v[i].z), __uint_as_float(v[i].w)));
          }
          flush(cur);
        }
      }
    }
    EW_PROF(15);
    if (p.debug_bar) asm volatile("bar.sync 1, 512;" ::: "memory");
  }

  // ---- RF: the band this warp stored in AG takes the next tile's P_j rows, once the store has READ it.
  // `newer`: bulk groups this thread committed after that store that may stay pending.  (Running RF one stage later,
  // behind the other context's AG, was measured: 197 us against 190 us -- the gather then lands too late for E0.)
  __device__ __forceinline__ void rf(int c, int t, int newer) {
    if (t + 1 < n_of(c)) {
      if (newer) tma::bulk_wait_read1();
      else tma::bulk_wait_read0();
      __syncwarp();
      load_pj(c, t + 1, slot(c, (uint32_t)t & 1u));
    }
    EW_PROF(16);
  }
};

template <bool BF, bool RELU_E, bool PROF, bool SAVE = false>
__global__ void __launch_bounds__(EW_THREADS, 1) in_edge_ws_kernel(const __grid_constant__ EwParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sm0 = smem_u32(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + EW_TMEM_SLOT);

  if (sm0 & 1023u) {  // swizzle atoms need the 1024-byte alignment the layout assumes
    if (tid == 0) atomicExch(&g_ew_fault, 3);
    __trap();
  }
  {  // packed weights and biases -> shared memory (generic proxy), visible to the tensor core after the proxy fence
    const float4* g4 = reinterpret_cast<const float4*>(p.packed);
    float4* s4 = reinterpret_cast<float4*>(smem_raw);
    for (int i = tid; i < EW_W / 16; i += EW_THREADS) s4[i] = __ldg(g4 + i);
    float4* b4 = reinterpret_cast<float4*>(smem_raw + EW_BIAS);
    for (int i = tid; i < 768 / 16; i += EW_THREADS) b4[i] = __ldg(g4 + EW_W / 16 + i);
  }
  if (tid == 0) {
    for (int c = 0; c < 2; ++c) {
      unsigned char* b = smem_raw + EW_BARS + 64 * c;
      mbar_init(reinterpret_cast<uint64_t*>(b + 0), 1);    // full_e: the producer's expect_tx
      mbar_init(reinterpret_cast<uint64_t*>(b + 8), 16);   // full_p: one expect_tx per owner warp
      mbar_init(reinterpret_cast<uint64_t*>(b + 16), 16);  // pj_free
      mbar_init(reinterpret_cast<uint64_t*>(b + 24), 16);  // a_ready
      mbar_init(reinterpret_cast<uint64_t*>(b + 32), 1);   // d_ready
      mbar_init(reinterpret_cast<uint64_t*>(b + 40), 16);  // out_ready
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

  // tiles of context c: blockIdx.x + gridDim.x * (2 t + c)
  const int g = (int)gridDim.x;
  int tile0[2], n_t[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    tile0[c] = (int)blockIdx.x + g * c;
    n_t[c] = tile0[c] < p.n_tiles ? (p.n_tiles - tile0[c] + 2 * g - 1) / (2 * g) : 0;
  }
  const bool prof_on = PROF && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 16 || warp == 18);
  long long prof_t = prof_on ? clock64() : 0;

  if (warp < 16) {
    // ================================================================= row owners
    EwOwner<BF, RELU_E, PROF, SAVE> o{p, sm0};
    o.w = warp; o.lane = lane; o.r = 32 * (warp & 3) + lane; o.qd = warp >> 2;
    o.rx = (uint32_t)(o.r & 7) << 4;
    o.own = (uint32_t)(o.qd >> 1) * 16384u + (uint32_t)o.r * 128u;
    o.tile00 = tile0[0]; o.nA = n_t[0]; o.nB = n_t[1];  // nB <= nA <= nB + 1
    o.prof_on = prof_on; o.prof_t = prof_t;
#pragma unroll 1
    for (int c = 0; c < 2; ++c)
      if (o.n_of(c) > 0) o.load_pj(c, 0, o.slot(c, 1));
#pragma unroll 1
    for (int c = 0; c < 2; ++c)
      if (o.n_of(c) > 0) o.c0(c, 0);
    // stage order of one round (two tiles): E0 E0' E1 E1' E2 E2' AG RF C0+ AG' RF' C0'+  -- a stage of one context runs
    // under the Linear of the other; RF (the P_j gather of the next tile) sits two stages in front of the E0 that needs it
#pragma unroll 1
    for (int t = 0; t < o.nA; ++t) {
#pragma unroll 1
      for (int c = 0; c < 2; ++c)
        if (t < o.n_of(c)) o.e0(c, t);
#pragma unroll 1
      for (int c = 0; c < 2; ++c)
        if (t < o.n_of(c)) o.e1(c, t);
#pragma unroll 1
      for (int c = 0; c < 2; ++c)
        if (t < o.n_of(c)) o.e2(c, t);
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        if (t < o.n_of(c)) {
          o.ag(c, t);
          o.rf(c, t, 0);
        }
        if (t + 1 < o.n_of(c)) o.c0(c, t + 1);
      }
      if (PROF && prof_on) g_ew_prof[31] += 1;
    }
    tma::bulk_wait_all0();
  } else if (warp < 18) {
    // ================================================================= TMA producer of context `ctx`
    const int ctx = warp - 16;
    const uint32_t bars = sm0 + EW_BARS + 64 * ctx;
    const uint32_t full_e = bars, pj_free = bars + 16;
    const uint32_t ids = sm0 + EW_IDS + 1024 * ctx;
    if (lane == 0) {
      tma::prefetch_map(&p.e_map);
      tma::prefetch_map(&p.pj_map);
      tma::prefetch_map(&p.out_map);
    }
    const bool e_gather = p.e_index != nullptr;
    for (int t = 0; t < n_t[ctx]; ++t) {
      const int tile = tile0[ctx] + t * 2 * g;
      const uint32_t row0 = (uint32_t)tile * EW_TM;
      const int rows_here = (int)min((int64_t)EW_TM, p.n_rows - (int64_t)row0);
      const uint32_t sl = sm0 + EW_SLOTS + (2 * ctx + (t & 1)) * EW_SLOT;
      EW_PROF(17);
      if (t > 0) ew_wait(pj_free, (uint32_t)(t - 1) & 1u);  // slot t & 1 held P_j(t - 1)
      EW_PROF(18);
      const bool full = rows_here == EW_TM;
      if (lane == 0) tma::mbar_expect_tx(full_e, EW_SLOT + (full ? 1024 : 0));
      __syncwarp();
      if (e_gather) {
        const int4 r4 = ew_idx4(p.e_index, row0, rows_here, lane, 0);
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
          const int a = __shfl_sync(0xffffffffu, r4.x, j), b = __shfl_sync(0xffffffffu, r4.y, j);
          const int c = __shfl_sync(0xffffffffu, r4.z, j), d = __shfl_sync(0xffffffffu, r4.w, j);
          if (elect_one()) {
            tma::gather4(sl + j * 512, &p.e_map, full_e, 0, a, b, c, d);
            tma::gather4(sl + 16384 + j * 512, &p.e_map, full_e, BF ? 64 : 32, a, b, c, d);
          }
          __syncwarp();
        }
      }
      if (elect_one()) {
        if (!e_gather) {
          tma::load_2d(sl, &p.e_map, full_e, 0, (int)row0);  // rows past the end of the table arrive as zeros
          tma::load_2d(sl + 16384, &p.e_map, full_e, BF ? 64 : 32, (int)row0);
        }
        if (full) {  // ids of the tile (a partial tile reads them with guarded loads instead)
          tma::bulk_g2s(ids, p.dst + row0, 512, full_e);
          tma::bulk_g2s(ids + 512, p.src + row0, 512, full_e);
        }
      }
      __syncwarp();
      EW_PROF(19);
    }
  } else {
    // ================================================================= MMA issue of context `ctx`
    const int ctx = warp - 18;
    const uint32_t bars = sm0 + EW_BARS + 64 * ctx;
    const uint32_t a_ready = bars + 24, d_ready = bars + 32;
    const uint32_t tmc = (uint32_t)ctx * EW_CTX;
    const uint32_t idesc = BF ? make_idesc_bf16(EW_TM, 128) : make_idesc_tf32(EW_TM, 64);
    int n = 0;  // commits so far
    for (int t = 0; t < n_t[ctx]; ++t) {
#pragma unroll 1
      for (int l = 0; l < 3; ++l, ++n) {
        EW_PROF(24);
        ew_wait(a_ready, (uint32_t)n & 1u);
        EW_PROF(25);
        tc_fence_after_sync();
        const uint64_t bd_hi = make_smem_desc_sw128(sm0 + (uint32_t)l * 32768u);
        const uint64_t bd_lo = make_smem_desc_sw128(sm0 + (uint32_t)l * 32768u + 16384u);
        if (elect_one()) {
          if constexpr (BF) {
            // 128 x 128 x 128 in bf16: eight K = 16 steps; the weights of a layer are two K tiles [128][64 bf16]
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_bf16_ts(tmc + 64u, tmc + 8 * ks, (ks < 4 ? bd_hi : bd_lo) + (uint64_t)((ks & 3) * 2), idesc, ks > 0);
          } else {
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
          }
          mma_commit_addr(d_ready);
        }
        __syncwarp();
        EW_PROF(26);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0u, 512);
}

// ------------------------------------------------------------------------------ host side
template <bool BF, bool RELU_E, bool PROF, bool SAVE = false>
static cudaError_t ew_configure() {
  return cudaFuncSetAttribute(in_edge_ws_kernel<BF, RELU_E, PROF, SAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, EW_SMEM);
}

static int ew_launch(EwParams& p, bool bf, bool relu, cudaStream_t st) {
  static const bool debug_bar = getenv("GTB_EW_DEBUG_BAR") != nullptr;
  p.debug_bar = debug_bar ? 1 : 0;
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = ew_configure<false, false, false>();
    if (e == cudaSuccess) e = ew_configure<false, true, false>();
    if (e == cudaSuccess) e = ew_configure<false, false, true>();
    if (e == cudaSuccess) e = ew_configure<false, true, true>();
    if (e == cudaSuccess) e = ew_configure<true, false, false>();
    if (e == cudaSuccess) e = ew_configure<true, true, false>();
    if (e == cudaSuccess) e = ew_configure<false, false, false, true>();
    if (e == cudaSuccess) e = ew_configure<false, true, false, true>();
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(in_edge_ws)");
    configured = true;
  }
  const int pairs = (p.n_tiles + 1) / 2;
  const int grid = pairs < kNumSMs ? pairs : kNumSMs;
  if (bf) {
    if (relu) in_edge_ws_kernel<true, true, false><<<grid, EW_THREADS, EW_SMEM, st>>>(p);
    else      in_edge_ws_kernel<true, false, false><<<grid, EW_THREADS, EW_SMEM, st>>>(p);
  } else if (p.h0 != nullptr) {
    if (relu) in_edge_ws_kernel<false, true, false, true><<<grid, EW_THREADS, EW_SMEM, st>>>(p);
    else      in_edge_ws_kernel<false, false, false, true><<<grid, EW_THREADS, EW_SMEM, st>>>(p);
  } else if (g_ew_prof_enabled) {
    if (relu) in_edge_ws_kernel<false, true, true><<<grid, EW_THREADS, EW_SMEM, st>>>(p);
    else      in_edge_ws_kernel<false, false, true><<<grid, EW_THREADS, EW_SMEM, st>>>(p);
  } else {
    if (relu) in_edge_ws_kernel<false, true, false><<<grid, EW_THREADS, EW_SMEM, st>>>(p);
    else      in_edge_ws_kernel<false, false, false><<<grid, EW_THREADS, EW_SMEM, st>>>(p);
  }
  GTB_CHECK_LAUNCH("in_edge_ws_kernel");
  return GTB_OK;
}

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

// The rows of the gathered tables are not part of the C descriptor; a TMA map only needs an upper bound for
// its bounds check, the indices are the plan's (clamped into their tables when the plan is built).  The
// projected block flagged SORTED (P_i) must be indexed by seg_id itself: one id array serves both.
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
      !tma::make_map_2d(&p.out_map, d.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, o_rows, 64, (uint64_t)d.out_ld, 32, o_scatter ? 1 : 8))
    return GTB_OK;  // the driver refused a map: the generic tiles take the launch
  p.pi = spi.ptr;
  p.pi_ld = spi.ld;
  p.dst = d.seg_id;
  p.src = spj.index;
  p.e_index = se.index;
  p.out_index = d.out_index;
  p.aggr = d.aggr;
  p.aggr_ld = d.aggr_ld;
  p.packed = static_cast<const unsigned char*>(d.packed);
  p.n_rows = d.n_rows;
  p.n_tiles = (int32_t)((d.n_rows + EW_TM - 1) / EW_TM);
  p.out = d.out;
  p.out_ld = d.out_ld;
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
  return ew_launch(p, false, se.relu != 0, st);
}

bool in_edge_ws_takes(const gtb_mlp_desc_t& d) {
  static const bool disabled = getenv("GTB_NO_EDGE_WS") != nullptr;
  int s_e, s_pi, s_pj;
  return !disabled && ew_match(d, &s_e, &s_pi, &s_pj) && tma::encode_fn() != nullptr;
}

// ------------------------------------------------------------------------------ bf16 variant (128 / 128 / 128)
// Packed image: per layer two K tiles [128 n][64 bf16] in the 128-byte-swizzle UMMA layout (16 KB each), then the
// three bias rows as bf16 (autocast casts the bias with the weight): 3 * 32 KB + 768 B, the fp32 image's size.
__global__ void pack_ew_bf16_kernel(const float* __restrict__ w0, const float* __restrict__ w1, const float* __restrict__ w2,
                                    const float* __restrict__ b0, const float* __restrict__ b1, const float* __restrict__ b2,
                                    unsigned char* __restrict__ packed) {
  const float* ws[3] = {w0, w1, w2};
  const float* bs[3] = {b0, b1, b2};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * 128 * 128 + 3 * 128; i += gridDim.x * blockDim.x) {
    if (i < 3 * 128 * 128) {
      const int l = i / 16384, n = (i >> 7) & 127, k = i & 127;
      const uint32_t kb = (uint32_t)(k & 63) * 2u;
      const uint32_t off = (uint32_t)l * 32768u + (uint32_t)(k >> 6) * 16384u + (uint32_t)(n >> 3) * 1024u + (uint32_t)(n & 7) * 128u +
                           ((((kb >> 4) ^ (uint32_t)(n & 7)) & 7u) << 4) + (kb & 15u);
      *reinterpret_cast<__nv_bfloat16*>(packed + off) = __float2bfloat16_rn(ws[l][n * 128 + k]);
    } else {
      const int j = i - 3 * 128 * 128, l = j >> 7, n = j & 127;
      reinterpret_cast<__nv_bfloat16*>(packed + EW_W)[j] = __float2bfloat16_rn(bs[l] ? bs[l][n] : 0.f);
    }
  }
}

int ew_pack_bf16(const float* const* weights, const float* const* biases, void* packed, cudaStream_t st) {
  GTB_REQUIRE(weights && weights[0] && weights[1] && weights[2] && packed, GTB_ERR_BAD_ARG, "gtb_in_edge_bf16_pack: null argument");
  pack_ew_bf16_kernel<<<96, 256, 0, st>>>(weights[0], weights[1], weights[2], biases ? biases[0] : nullptr,
                                          biases ? biases[1] : nullptr, biases ? biases[2] : nullptr,
                                          static_cast<unsigned char*>(packed));
  GTB_CHECK_LAUNCH("pack_ew_bf16_kernel");
  return GTB_OK;
}

int in_edge_ws_bf16(const void* e_in, int32_t e_ld, const int32_t* e_index, int32_t relu_e, const void* p_i, int32_t pi_ld,
                    const void* p_j, int32_t pj_ld, int64_t n_edges, const int32_t* src_sorted, const int32_t* dst_sorted,
                    const void* packed, void* e_out, int32_t eo_ld, const int32_t* out_index, float* aggr, int32_t aggr_ld,
                    cudaStream_t st) {
  GTB_REQUIRE(e_in && p_i && p_j && src_sorted && dst_sorted && packed && e_out && aggr, GTB_ERR_BAD_ARG,
              "gtb_in_edge_forward_bf16: null argument");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  GTB_REQUIRE(al16(e_in) && al16(p_i) && al16(p_j) && al16(e_out) && al16(aggr) && al16(src_sorted) && al16(dst_sorted) &&
                  al16(packed) && (e_index == nullptr || al16(e_index)) && (out_index == nullptr || al16(out_index)),
              GTB_ERR_BAD_ARG, "gtb_in_edge_forward_bf16: pointers must be 16-byte aligned");
  GTB_REQUIRE(e_ld >= 128 && pi_ld >= 128 && pj_ld >= 128 && eo_ld >= 128 && aggr_ld >= 128 && !(e_ld & 7) && !(pi_ld & 7) &&
                  !(pj_ld & 7) && !(eo_ld & 7) && !(aggr_ld & 3),
              GTB_ERR_BAD_ARG, "gtb_in_edge_forward_bf16: row strides must cover 128 columns in 16-byte steps");
  GTB_REQUIRE(n_edges >= 0 && n_edges < (1ll << 31) - 256, GTB_ERR_BAD_ARG, "gtb_in_edge_forward_bf16: bad n_edges");
  GTB_REQUIRE(tma::encode_fn() != nullptr, GTB_ERR_CUDA, "gtb_in_edge_forward_bf16: cuTensorMapEncodeTiled is not available");
  if (n_edges == 0) return GTB_OK;
  EwParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t big = 1ull << 31;
  const bool e_gather = e_index != nullptr, o_scatter = out_index != nullptr;
  const uint64_t e_rows = e_gather ? big : (uint64_t)n_edges, o_rows = o_scatter ? big : (uint64_t)n_edges;
  const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  GTB_REQUIRE(tma::make_map_2d(&p.e_map, e_in, bf, 2, e_rows, 128, (uint64_t)e_ld, 64, e_gather ? 1 : 128) &&
                  tma::make_map_2d(&p.pj_map, p_j, bf, 2, big, 128, (uint64_t)pj_ld, 64, 1) &&
                  tma::make_map_2d(&p.out_map, e_out, bf, 2, o_rows, 128, (uint64_t)eo_ld, 64, o_scatter ? 1 : 8),
              GTB_ERR_CUDA, "gtb_in_edge_forward_bf16: the driver refused a tensor map");
  p.pi = p_i;
  p.pi_ld = pi_ld;
  p.dst = dst_sorted;
  p.src = src_sorted;
  p.e_index = e_index;
  p.out_index = out_index;
  p.aggr = aggr;
  p.aggr_ld = aggr_ld;
  p.packed = static_cast<const unsigned char*>(packed);
  p.n_rows = n_edges;
  p.n_tiles = (int32_t)((n_edges + EW_TM - 1) / EW_TM);
  p.out = e_out;
  p.out_ld = eo_ld;
  return ew_launch(p, true, relu_e != 0, st);
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
