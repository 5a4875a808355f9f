// Fused row MLP on the 5th-generation tensor cores (tcgen05, sm_100a): fp32 in / fp32 out with
// the 3xTF32 split (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM), which holds the 1e-5
// parity bar that plain TF32 cannot (profiles/r1_tc_unit_probe.log: kind::tf32 truncates).
//
// One persistent CTA per SM walks 128-row tiles:
//   * rows of the streamed column blocks (edge features, node features) are gathered with
//     16-byte cp.async copies into XOR-swizzled shared-memory slots, a ring of slots keeping the
//     next items in flight while the current tile is computed;
//   * the row-owner thread splits its row into tf32 hi / lo parts and writes them to TMEM
//     (tcgen05.st); the A operand of every MMA comes from TMEM, the B operand (packed weights,
//     resident for the whole kernel) from 128-byte-swizzled shared memory;
//   * one thread issues the tcgen05.mma chain, completion arrives on an mbarrier
//     (tcgen05.commit); the epilogue reads the accumulator back (tcgen05.ld), adds the bias and
//     the gathered pre-projected rows, applies ReLU and feeds the next layer through TMEM again:
//     activations never touch shared or global memory between the Linear layers;
//   * the last epilogue stages the output tile in a slot: coalesced row stores (optionally
//     scattered by out_index, with the residual fused in) and the in-tile segmented sum by
//     destination.
// Reference semantics: see include/gtb200.h (gtb_fused_mlp_f32).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gtb {

using namespace tc;

constexpr int TC_TM = 128;                 // rows per tile = UMMA M
constexpr int TC_SLOT = TC_TM * 256;       // one staging slot: [128 rows][64 fp32], 16-byte chunks ^ (row & 7)
constexpr int TC_MAXCH = 8;                // streamed sub-blocks (<= 64 columns each) of the first Linear
constexpr int TC_MAXITEMS = TC_MAXCH + 3;
constexpr int TC_SMEM_MAX = 232448;        // 227 KB opt-in shared memory per CTA
constexpr int TC_HEAD = 1072;              // per team: 128 segment ids + two MMA barriers; TMEM slot
constexpr int TC_MISC = TC_HEAD + 1023;    // + slack to align the weight tiles to 1024 bytes
constexpr uint32_t TM_A_HI = 0, TM_A_LO = 64, TM_D = 128, TM_CTX = 192;  // TMEM column map of one team

constexpr int TC_IDX_IDENTITY = -1, TC_IDX_GLOBAL = -2;

__device__ int g_tc_timeout = 0;
__device__ long long g_tc_prof[32];
static int g_prof_enabled = 0;

// packed weights of one MLP: per layer the tf32 hi and lo parts as K-major [npad][32] fp32 tiles
// (128-byte swizzle), then one 64-float bias row per layer.  Layer 0 keeps each streamed column
// block padded to a multiple of 8 columns.
struct TcLayout {
  int n_layers, n_chunks;
  int chunk_w[GTB_MAX_SRCS], chunk_koff[GTB_MAX_SRCS], chunk_kpad[GTB_MAX_SRCS];
  int kpad[GTB_MAX_LAYERS], npad[GTB_MAX_LAYERS], ntrue[GTB_MAX_LAYERS], ktiles[GTB_MAX_LAYERS];
  uint32_t w_off[GTB_MAX_LAYERS][2], b_off[GTB_MAX_LAYERS], total_bytes;
  // a last Linear with <= 4 outputs behind a hidden layer (edge weights, beta) is evaluated on the
  // CUDA cores from the fp32 activations the row owners already hold: plain fp32 rows [4][kpad],
  // no MMA, no tf32 split of the last hidden activation
  int narrow_last;
};

struct TcChunk {
  const float* ptr;
  const int32_t* index;
  int32_t ld, width, kpad, koff, relu, staged;
};
struct TcAdd {
  const float* ptr;
  const int32_t* index;
  int32_t ld, staged;
};
struct TcParams {
  int64_t n_rows;
  int32_t n_tiles, n_chunks, n_adds, n_layers, ring, ipt, n_teams;
  int8_t items[TC_MAXITEMS + 1];  // per tile, in consumption order: chunk c -> c, add a -> 64 + a, output tile -> -1
  // row indices of the gather copies travel in registers, fetched two tiles ahead by the lanes that
  // will use them (no exposed index latency, no shared memory): per item the register slot
  // (TC_IDX_IDENTITY: rows are the tile's own rows, TC_IDX_GLOBAL: irregular width, read on use)
  int8_t item_ireg[TC_MAXITEMS + 1];
  int32_t n_iregs, ireg_c4n[4];
  const int32_t* ireg_ptr[4];
  int32_t out_mode, out_c4n;      // same for the output rows (out_index)
  int32_t prof;                   // debug: per-stage clock accumulation by one thread
  int32_t narrow_last;            // see TcLayout
  TcChunk ch[TC_MAXCH];
  TcAdd add[2];
  int32_t kpad[GTB_MAX_LAYERS], npad[GTB_MAX_LAYERS], ntrue[GTB_MAX_LAYERS];
  uint32_t w_off[GTB_MAX_LAYERS][2], b_off[GTB_MAX_LAYERS], w_bytes;
  const unsigned char* packed;
  int32_t final_act;
  float act_eps, res_a, res_b;
  const float* res;
  int32_t res_ld;
  const float* row_scale;
  const float* out_scale;
  const float* gate;
  int32_t gate_ld;
  float* out;
  const int32_t* out_index;
  int32_t out_ld;
  float* aggr;
  int32_t aggr_ld;
  const int32_t* seg_id;
  const int32_t* rowptr;
};

// ------------------------------------------------------------------------------ layout
bool tc_layout(int n_layers, const int32_t* dims, int n_chunks, const int32_t* chunk_w, TcLayout* L) {
  if (n_layers < 1 || n_layers > GTB_MAX_LAYERS) return false;
  int32_t one = dims[0];
  if (n_chunks <= 0 || chunk_w == nullptr) {
    n_chunks = 1;
    chunk_w = &one;
  }
  if (n_chunks > GTB_MAX_SRCS) return false;
  int k0 = 0, sum = 0;
  L->n_chunks = n_chunks;
  for (int c = 0; c < n_chunks; ++c) {
    const int w = chunk_w[c];
    if (w < 1) return false;
    // a block is either staged through shared memory in 16-byte pieces or, when narrow, read
    // directly by the row-owner thread
    if (w % 4 != 0 && w > 16) return false;
    L->chunk_w[c] = w;
    L->chunk_koff[c] = k0;
    L->chunk_kpad[c] = round_up(w, 8);
    k0 += L->chunk_kpad[c];
    sum += w;
  }
  if (sum != dims[0]) return false;
  L->n_layers = n_layers;
  L->narrow_last = n_layers >= 2 && dims[n_layers] <= 4;
  uint32_t off = 0;
  for (int l = 0; l < n_layers; ++l) {
    const int n = dims[l + 1];
    if (n < 1 || n > 64) return false;
    L->ntrue[l] = n;
    L->npad[l] = round_up(n, 16);
    L->kpad[l] = (l == 0) ? k0 : L->npad[l - 1];
    L->ktiles[l] = (L->kpad[l] + 31) / 32;
    if (l == n_layers - 1 && L->narrow_last) {
      L->w_off[l][0] = L->w_off[l][1] = off;
      off += (uint32_t)n * (uint32_t)L->kpad[l] * 4u;  // n rows of kpad floats: a multiple of 64 bytes
      continue;
    }
    const uint32_t bytes = (uint32_t)L->ktiles[l] * L->npad[l] * 128u;  // multiple of 2048
    L->w_off[l][0] = off;
    off += bytes;
    L->w_off[l][1] = off;
    off += bytes;
  }
  for (int l = 0; l < n_layers; ++l) {
    L->b_off[l] = off;
    off += (l == n_layers - 1 && L->narrow_last) ? 4 * 4 : 64 * 4;
  }
  L->total_bytes = off;
  return (size_t)off + TC_SLOT + TC_MISC <= (size_t)TC_SMEM_MAX;
}

// staging slots (32 KB each) left beside the packed weights: >= 4 lets two teams share an SM,
// 1 leaves every gather latency exposed; 0: the widths are not supported
int tc_slots(int n_layers, const int32_t* dims, int n_chunks, const int32_t* chunk_w) {
  TcLayout L;
  if (!tc_layout(n_layers, dims, n_chunks, chunk_w, &L)) return 0;
  return (int)((TC_SMEM_MAX - (((size_t)L.total_bytes + 15) / 16 * 16 + TC_MISC)) / TC_SLOT);
}

size_t tc_packed_bytes(int n_layers, const int32_t* dims, int n_chunks, const int32_t* chunk_w) {
  TcLayout L;
  if (!tc_layout(n_layers, dims, n_chunks, chunk_w, &L)) return 0;
  return L.total_bytes;
}

struct PackArgs {
  TcLayout L;
  const float* w[GTB_MAX_LAYERS];
  const float* b[GTB_MAX_LAYERS];
  int32_t ktrue[GTB_MAX_LAYERS];
};

// One thread per padded (layer, n, k): tf32 hi / lo parts into the swizzled K-major tiles.
__global__ void pack_tc_kernel(const __grid_constant__ PackArgs a, unsigned char* __restrict__ packed) {
  const TcLayout& L = a.L;
  for (int l = 0; l < L.n_layers; ++l) {
    if (l == L.n_layers - 1 && L.narrow_last) {  // plain fp32 rows [4][kpad] + bias
      const int kp = L.kpad[l];
      const int nr = L.ntrue[l];
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nr * kp + 4; i += gridDim.x * blockDim.x) {
        if (i < nr * kp) {
          const int n = i / kp, k = i - n * kp;
          reinterpret_cast<float*>(packed + L.w_off[l][0])[i] = k < a.ktrue[l] ? a.w[l][(size_t)n * a.ktrue[l] + k] : 0.f;
        } else {
          const int n = i - nr * kp;
          reinterpret_cast<float*>(packed + L.b_off[l])[n] = (a.b[l] != nullptr && n < L.ntrue[l]) ? a.b[l][n] : 0.f;
        }
      }
      continue;
    }
    const int kext = L.ktiles[l] * 32, npad = L.npad[l];
    const int total = kext * npad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total + npad; i += gridDim.x * blockDim.x) {
      if (i >= total) {
        const int n = i - total;
        reinterpret_cast<float*>(packed + L.b_off[l])[n] = (a.b[l] != nullptr && n < L.ntrue[l]) ? a.b[l][n] : 0.f;
        continue;
      }
      const int n = i / kext, k = i - n * kext;
      float v = 0.f;
      if (n < L.ntrue[l]) {
        if (l == 0) {
          int col = 0;
          for (int c = 0; c < L.n_chunks; ++c) {
            const int j = k - L.chunk_koff[c];
            if (j >= 0 && j < L.chunk_w[c]) v = a.w[0][(size_t)n * a.ktrue[0] + col + j];
            col += L.chunk_w[c];
          }
        } else if (k < a.ktrue[l]) {
          v = a.w[l][(size_t)n * a.ktrue[l] + k];
        }
      }
      float hi, lo;
      split_tf32(v, hi, lo);
      const uint32_t o = (uint32_t)(k >> 5) * (uint32_t)npad * 128u + sw128_offset(n, k & 31);
      *reinterpret_cast<float*>(packed + L.w_off[l][0] + o) = hi;
      *reinterpret_cast<float*>(packed + L.w_off[l][1] + o) = lo;
    }
  }
}

int pack_tc(int n_layers, const int32_t* dims, int n_chunks, const int32_t* chunk_w, const float* const* weights,
            const float* const* biases, void* packed, cudaStream_t st) {
  PackArgs a;
  memset(&a, 0, sizeof(a));
  GTB_REQUIRE(tc_layout(n_layers, dims, n_chunks, chunk_w, &a.L), GTB_ERR_UNSUPPORTED_DIM,
              "gtb_mlp_pack: these widths are not supported by the tcgen05 path");
  for (int l = 0; l < n_layers; ++l) {
    a.w[l] = weights[l];
    a.b[l] = biases ? biases[l] : nullptr;
    a.ktrue[l] = dims[l];
  }
  pack_tc_kernel<<<64, 256, 0, st>>>(a, static_cast<unsigned char*>(packed));
  GTB_CHECK_LAUNCH("pack_tc_kernel");
  return GTB_OK;
}

// ------------------------------------------------------------------------------ kernel
// A CTA runs one or two TEAMS of 256 threads.  A team owns its tiles (alternating with the other
// team), its ring of staging slots, its TMEM columns and its MMA barrier, and synchronises on a
// named barrier only -- so one team's MMA chain and load latency overlap the other team's
// epilogue / conversion work on the same SM, with the packed weights shared.
//
// Code-generation notes (measured with ncu, profiles/README.md): the tile loop is issue-bound, so
// shared memory is addressed with 32-bit shared-window addresses (no generic-pointer arithmetic),
// ring positions are counted, never divided, and row offsets are 32 x 32 -> 64-bit multiplies.
constexpr int TC_TEAM = 256;

__device__ __forceinline__ void team_sync(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(TC_TEAM) : "memory");
}
__device__ __forceinline__ uint32_t slot_off(int r, int c4) {  // 16-byte chunk c4 of row r
  return (uint32_t)(r * 256 + ((c4 ^ (r & 7)) << 4));
}

__device__ __forceinline__ void wait_or_trap(uint32_t bar_addr, uint32_t parity) {
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
        : "r"(bar_addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  atomicExch(&g_tc_timeout, 1);
  __trap();  // a wrong descriptor must fail loudly, never hang the GPU or return garbage
}

__device__ __forceinline__ int team_tile(const TcParams& p, int team, int t) {
  return (int)blockIdx.x + (int)gridDim.x * (t * p.n_teams + team);
}

// Row (inside the tile) whose index copy slot `lane` of warp `wteam` holds, for a block of
// c4n = 2^sh 16-byte pieces per row, or -1: thread tt = 32 wteam + lane issues the copies
// i = tt + 256 j (row i >> sh), so a warp touches the rows ((32 wteam + 256 j) >> sh) + q,
// q < m = 32 >> sh, j < J = max(1, c4n / 2): 16 rows (32 for c4n = 1); lane s = j m + q.
__device__ __forceinline__ int tc_copy_slot_row(int c4n, int wteam, int lane) {
  const int sh = __ffs(c4n) - 1;
  const int m = 32 >> sh;
  const int J = c4n >= 2 ? (c4n >> 1) : 1;
  const int j = lane >> (5 - sh), q = lane & (m - 1);
  return j < J ? ((32 * wteam + 256 * j) >> sh) + q : -1;
}

// gather one staged item (kind: streamed block c -> c, pre-projected block a -> 64 + a, output
// tile -> -1) of team tile `tile` into the slot at shared address `sdst`; every thread commits
// exactly one cp.async group per call.  idxv: this warp's row-index register for that tile.
// W64: the block is 64 columns wide: thread tt copies piece (tt & 15) of the rows (tt >> 4) + 16 j,
// whose slot addresses differ by a constant 4096 bytes.
template <bool W64>
__device__ __forceinline__ void tc_issue_item(const TcParams& p, int tile, int k, uint32_t sdst, int tt, int32_t idxv) {
  const int kind = p.items[k];
  if (tile < p.n_tiles && kind >= 0) {
    const float* ptr;
    const int32_t* index;
    uint32_t ld4;
    int width;
    if (kind < 64) {
      ptr = p.ch[kind].ptr; index = p.ch[kind].index; ld4 = (uint32_t)p.ch[kind].ld * 4u; width = p.ch[kind].width;
    } else {
      ptr = p.add[kind - 64].ptr; index = p.add[kind - 64].index; ld4 = (uint32_t)p.add[kind - 64].ld * 4u; width = p.ntrue[0];
    }
    const int mode = p.item_ireg[k];
    const uint32_t row0 = (uint32_t)tile * TC_TM;
    const int rows_here = (int)min((int64_t)TC_TM, p.n_rows - (int64_t)row0);
    const int lane = tt & 31;
    if (W64) {
      const int rb = tt >> 4, c = tt & 15;
      const uint32_t d0 = sdst + slot_off(rb, c);  // rows rb + 16 j share (row & 7): same swizzle
      const float* src0 = ptr + (c << 2);
      if (mode >= 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t row = (uint32_t)__shfl_sync(0xffffffffu, idxv, 2 * j + (lane >> 4));
          if (rb + 16 * j < rows_here) cp_async16(d0 + j * 4096, row_ptr(src0, row, ld4));
        }
      } else if (mode == TC_IDX_IDENTITY) {
        const float* s = row_ptr(src0, row0 + rb, ld4);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (rb + 16 * j < rows_here) cp_async16(d0 + j * 4096, reinterpret_cast<const char*>(s) + (uint64_t)(16 * j) * ld4);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (rb + 16 * j < rows_here)
            cp_async16(d0 + j * 4096, row_ptr(src0, (uint32_t)__ldg(index + row0 + rb + 16 * j), ld4));
      }
    } else {
      const int c4n = width >> 2;
      const int total = rows_here * c4n;  // <= 8 copies per thread
      const bool pow2 = (c4n & (c4n - 1)) == 0;
      const int sh = __ffs(c4n) - 1;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = tt + j * TC_TEAM;
        const int r = pow2 ? (i >> sh) : (i / c4n);
        uint32_t row = row0 + r;
        if (mode >= 0) row = (uint32_t)__shfl_sync(0xffffffffu, idxv, (j * (32 >> sh) + (lane >> sh)) & 31);
        else if (mode == TC_IDX_GLOBAL && i < total) row = (uint32_t)__ldg(index + row0 + r);
        if (i < total) {
          const int c = i - r * c4n;
          cp_async16(sdst + slot_off(r, c), row_ptr(ptr + (c << 2), row, ld4));
        }
      }
    }
  }
  cp_async_commit();
}

// MMA issue.  The whole first warp of a team calls these with warp-uniform arguments; ONE lane,
// chosen by elect.sync, runs the branch that holds the tcgen05.mma chain.  ptxas recognises the
// elect-guarded branch as single-lane code and keeps descriptors / TMEM addresses in uniform
// registers: back-to-back UTCHMMA with UIADD3 in between.  The same chain issued from
// `if (tid == 0)` compiles to an ELECT / 6 x R2UR / UTCHMMA / BRA.U.ANY waterfall loop per MMA
// (~150 cycles each, measured: profiles/r1_tc_mma_bench.log), 5x the 32 cycles the tensor core
// needs for M = 128, N = 64, K = 8.  The TMEM base is 0 (checked at kernel start).
// three passes (small terms first: lo*hi, hi*lo, hi*hi) over `ksteps` K = 8 steps of one streamed
// block (first Linear) or of a whole hidden Linear, then the commit onto the team's barrier
template <int TEAM>
__device__ __forceinline__ void tc_issue_mmas(uint32_t wbase, uint32_t bar_base, const TcParams& p, int l, int koff,
                                              int ksteps, bool first) {
  constexpr uint32_t tmc = TEAM * TM_CTX;
  const uint32_t idesc = make_idesc_tf32(TC_TM, p.npad[l]);
  const uint32_t tile16 = (uint32_t)p.npad[l] * 8u;  // bytes of one [npad][32] K tile >> 4
  const uint32_t boff = (uint32_t)(koff >> 5) * tile16 + (uint32_t)((koff & 31) >> 3) * 2u;
  const uint64_t bd_hi = make_smem_desc_sw128(wbase + p.w_off[l][0]) + boff;
  const uint64_t bd_lo = make_smem_desc_sw128(wbase + p.w_off[l][1]) + boff;
  if (elect_one()) {
    bool acc = !first;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
      uint32_t a = tmc + ((pass == 0) ? TM_A_LO : TM_A_HI);
      uint64_t bd = (pass == 1) ? bd_lo : bd_hi;
      int sub = (koff & 31) >> 3;
#pragma unroll 1
      for (int ks = 0; ks < ksteps; ++ks) {
        mma_tf32_ts(tmc + TM_D, a, bd, idesc, acc);
        acc = true;
        a += 8;
        if (++sub == 4) {
          sub = 0;
          bd += tile16 - 6;  // next 32-wide K tile
        } else {
          bd += 2;           // +32 bytes inside the swizzled tile
        }
      }
    }
    mma_commit_addr(bar_base + TEAM * 16);
  }
  __syncwarp();
}

// The common case -- a 64-wide block / layer starting at K = 0 -- with the layer index a template
// constant and everything unrolled: operands are kernel-parameter loads plus immediates, which
// ptxas keeps in uniform registers (no R2UR in front of every UTCHMMA).
// `halves`: a 64-wide output is issued as two N = 32 chains with one commit each, output columns
// 0..31 first: the epilogue of the first half then runs under the MMAs of the second.
// `ktile`: first 32-wide K tile of the weights this 64-wide block multiplies (koff / 32).
template <int TEAM, int L>
__device__ __forceinline__ void tc_issue_mmas_k64(uint32_t wbase, uint32_t bar_base, const TcParams& p, bool first,
                                                  bool halves, uint32_t ktile = 0) {
  constexpr uint32_t tmc = TEAM * TM_CTX;
  const uint32_t n = halves ? 32u : (uint32_t)p.npad[L];
  const uint32_t idesc = make_idesc_tf32(TC_TM, (int)n);
  const uint32_t tile16 = (uint32_t)p.npad[L] * 8u;
  const uint64_t bd_hi = make_smem_desc_sw128(wbase + p.w_off[L][0]) + (uint64_t)(ktile * tile16);
  const uint64_t bd_lo = make_smem_desc_sw128(wbase + p.w_off[L][1]) + (uint64_t)(ktile * tile16);
  if (elect_one()) {  // ptxas knows a single lane runs this branch: operands go to uniform registers once
#pragma unroll 1
    for (uint32_t hf = 0; hf < (halves ? 2u : 1u); ++hf) {
      bool acc = !first;
      const uint32_t d = tmc + TM_D + 32u * hf;
      const uint64_t rows = (uint64_t)(hf * 256u);  // 32 rows x 128 bytes of every K tile, >> 4
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = tmc + ((pass == 0) ? TM_A_LO : TM_A_HI);
        const uint64_t bd = ((pass == 1) ? bd_lo : bd_hi) + rows;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          mma_tf32_ts(d, a + 8 * ks, bd + (uint64_t)((ks >> 2) * tile16 + (ks & 3) * 2), idesc, acc);
          acc = true;
        }
      }
      mma_commit_addr(bar_base + TEAM * 16 + hf * 8);  // one barrier per half: a waiter never lags two phases
    }
  }
  __syncwarp();
}

#define TC_PROF(id)                                   \
  do {                                                \
    if (PROF && prof_on) {                            \
      const long long now_ = clock64();               \
      g_tc_prof[id] += now_ - prof_t;                 \
      prof_t = now_;                                  \
    }                                                 \
  } while (0)

// W64: every staged streamed block, every hidden Linear and every pre-projected block is exactly 64
// wide (the "wide" Interaction-Network configuration; narrow blocks read directly by the row owner,
// as in the encoders, are fine): straight-line tile code without width guards.
// PROF: per-stage clock accumulation by thread 0 of CTA 0 (tests/cuda/tc_diag.py).
template <bool W64, bool PROF>
__global__ void __launch_bounds__(2 * TC_TEAM, 1) fused_mlp_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sm0 = smem_u32(smem_raw);
  // [2][128] segment ids | [2] MMA barriers | TMEM slot | (pad to 1024) weights | slot rings
  const uint32_t wbase = (sm0 + TC_HEAD + 1023u) & ~1023u;
  unsigned char* wsm = smem_raw + (wbase - sm0);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int team = tid >> 8, tt = tid & (TC_TEAM - 1), wteam = tt >> 5;
  const int r = tt & (TC_TM - 1), h = tt >> 7;
  const uint32_t ring = (uint32_t)p.ring;
  const uint32_t slots = wbase + ((p.w_bytes + 15u) & ~15u) + (uint32_t)team * ring * TC_SLOT;  // this team's ring
  const uint32_t segs = sm0 + (uint32_t)team * (TC_TM * 4);
  const uint32_t bar_base = sm0 + 2 * TC_TM * 4;
  const uint32_t mma_bar = bar_base + team * 16, mma_bar1 = mma_bar + 8;  // second barrier: second output half
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 2 * TC_TM * 4 + 32);
  const uint32_t rsw = (uint32_t)r * 256u, rx = (uint32_t)(r & 7) << 4;  // own row in a slot: rsw + ((c4 << 4) ^ rx)
  const bool prof_on = PROF && blockIdx.x == 0 && tid == 0;
  long long prof_t = prof_on ? clock64() : 0;

  // ---- row indices: copy-type (per warp slot, see tc_copy_slot_row) and row-type (per row owner;
  // arrays: 0 = segment ids, 1 / 2 = directly read pre-projected blocks), the first two tiles now, then
  // always two tiles ahead (see load_idx)
  const int32_t* rarr[3] = {p.seg_id, (p.n_adds > 0 && !p.add[0].staged) ? p.add[0].index : nullptr,
                            (p.n_adds > 1 && !p.add[1].staged) ? p.add[1].index : nullptr};
  int crow[4], orow_slot = -1;
#pragma unroll
  for (int q = 0; q < 4; ++q) crow[q] = q < p.n_iregs ? tc_copy_slot_row(p.ireg_c4n[q], wteam, lane) : -1;
  if (p.out_mode == 1) orow_slot = tc_copy_slot_row(p.out_c4n, wteam, lane);
  // Three generations of index registers: the tile being computed (cur), the tile whose copies are
  // being issued (next) and the tile after it (far), loaded at the top of a tile and first read a whole
  // tile later: ptxas parks the scoreboard wait of a load at the next branch, so a load read within
  // the same tile costs its full latency there (it did: ~10 % of the kernel, profiles/).
  int32_t ccur[4], cnext[4], cfar[4], rcur[3], rnext[3], rfar[3], ocur = 0, onext = 0;
  auto load_idx = [&](int tile_i, int32_t (&cdst)[4], int32_t (&rdst)[3], int32_t& odst) {
    const bool have = tile_i < p.n_tiles;
    const uint32_t row0i = have ? (uint32_t)tile_i * TC_TM : 0u;
    const int rows_i = have ? (int)min((int64_t)TC_TM, p.n_rows - (int64_t)row0i) : 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int32_t v = 0;
      if (crow[q] >= 0 && crow[q] < rows_i) v = __ldg(p.ireg_ptr[q] + row0i + crow[q]);
      cdst[q] = v;
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      int32_t v = 0;
      if (rarr[q] && r < rows_i) v = __ldg(rarr[q] + row0i + r);
      rdst[q] = v;
    }
    int32_t o = 0;
    if (orow_slot >= 0 && orow_slot < rows_i) o = __ldg(p.out_index + row0i + orow_slot);
    odst = o;
  };
  load_idx(team_tile(p, team, 0), ccur, rcur, ocur);
  load_idx(team_tile(p, team, 1), cnext, rnext, onext);
  // ring bookkeeping without divisions: next item to issue (tile sequence, kind position, slot)
  int iss_t = 0, iss_k = 0, issued = 0, consumed = 0;
  uint32_t iss_slot = 0, cons_slot = 0;
  auto issue_next = [&](int t_now) {
    const int ir = p.item_ireg[iss_k];
    int32_t idxv = 0;
    if (ir >= 0) {
      const bool nx = iss_t != t_now;
      const int32_t a0 = nx ? cnext[0] : ccur[0], a1 = nx ? cnext[1] : ccur[1];
      const int32_t a2 = nx ? cnext[2] : ccur[2], a3 = nx ? cnext[3] : ccur[3];
      idxv = ir == 0 ? a0 : ir == 1 ? a1 : ir == 2 ? a2 : a3;
    }
    tc_issue_item<W64>(p, team_tile(p, team, iss_t), iss_k, slots + iss_slot * TC_SLOT, tt, idxv);
    ++issued;
    if (++iss_slot == ring) iss_slot = 0;
    if (++iss_k == p.ipt) {
      iss_k = 0;
      ++iss_t;
    }
  };
  auto consume_done = [&]() {  // the item in cons_slot has been read by every thread of the team
    ++consumed;
    if (++cons_slot == ring) cons_slot = 0;
  };

  // ---- prologue: first items in flight, weights into shared memory, barriers + TMEM set-up
  for (int i = 0; i < p.ring; ++i) issue_next(0);
  {
    const float4* g4 = reinterpret_cast<const float4*>(p.packed);
    float4* s4 = reinterpret_cast<float4*>(wsm);
    for (int i = tid; i < (int)(p.w_bytes >> 4); i += blockDim.x) s4[i] = __ldg(g4 + i);
  }
  if (tt == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mma_bar), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mma_bar1), "r"(1) : "memory");
    fence_barrier_init();
  }
  const uint32_t tmem_cols = p.n_teams == 2 ? 512u : 256u;
  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  fence_proxy_async_smem();  // weights were written through the generic proxy, the MMA reads them through the async proxy
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = *tmem_slot;
  if (tm != 0) {  // one CTA per SM and one allocation: the base is column 0 / lane 0; the MMA issue relies on it
    atomicExch(&g_tc_timeout, 2);
    __trap();
  }
  const uint32_t tmc = (uint32_t)team * TM_CTX;                             // this team's TMEM columns
  const uint32_t tm_lane = tmc + ((uint32_t)((warp & 3) * 32) << 16);      // + this warp's 32 lanes
  uint32_t mma_phase = 0, mma_phase1 = 0;
  const int last = p.n_layers - 1;
  // A 64-wide LAST Linear of the straight-line variant is produced in two halves (tc_issue_mmas_k64):
  // its epilogue only writes shared memory.  Hidden layers are not: their epilogue overwrites the TMEM
  // A operand the second half's MMAs would still be reading.
  const bool halves0 = W64 && last == 0 && p.n_chunks == 1 && p.ch[0].staged && p.npad[0] == 64;
  TC_PROF(0);

  for (int t = 0;; ++t) {
    const int tile = team_tile(p, team, t);
    if (tile >= p.n_tiles) break;
    const uint32_t row0 = (uint32_t)tile * TC_TM;
    const int rows_here = (int)min((int64_t)TC_TM, p.n_rows - (int64_t)row0);
    const bool live = r < rows_here;
    if (tt < TC_TM) sts_i32(segs + tt * 4, (live && p.seg_id) ? rcur[0] : -1);
    // Indices of the tile after the next one: issued right in front of the first accumulator wait
    // of the tile (ptxas drains the scoreboard of a load at the next branch it meets, whatever the
    // distance to the first use; there the drain runs under the MMAs), first read a tile from now.
    int32_t ofar = 0;
    auto load_far = [&]() { load_idx(team_tile(p, team, t + 2), cfar, rfar, ofar); };
    const float rscale = (p.row_scale && live) ? __ldg(p.row_scale + row0 + r) : 1.f;

    // ---------------- first Linear: one streamed block at a time through the TMEM A buffer
    for (int c = 0; c < p.n_chunks; ++c) {
      const TcChunk& ch = p.ch[c];
      const int groups = ch.kpad >> 3;
      if (ch.staged) {
        cp_async_wait_pending(issued - consumed - 1);
        team_sync(team);
      }
      if (c > 0) {  // the previous block's MMAs still read the A buffer
        wait_or_trap(mma_bar, mma_phase);
        mma_phase ^= 1;
        tc_fence_after_sync();
      }
      TC_PROF(1);
      if (ch.staged) {
        const uint32_t sl = slots + cons_slot * TC_SLOT + rsw;
        const float clamp = ch.relu ? 0.f : -INFINITY;  // relu on load (layers > 0 of a residual stack)
        float4 a[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c4 = 2 * (h + 2 * q);
          a[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          b[q] = a[q];
          if (W64 || c4 * 4 < ch.width) a[q] = lds128(sl + (((uint32_t)c4 << 4) ^ rx));
          if (W64 || (c4 + 1) * 4 < ch.width) b[q] = lds128(sl + (((uint32_t)(c4 + 1) << 4) ^ rx));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int g8 = h + 2 * q;
          if (W64 || g8 < groups) {
            float v[8] = {a[q].x, a[q].y, a[q].z, a[q].w, b[q].x, b[q].y, b[q].z, b[q].w};
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], clamp);
            if (p.row_scale) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] *= rscale;
            }
            split_store8(tm_lane + TM_A_HI + 8 * g8, tm_lane + TM_A_LO + 8 * g8, v);
          }
        }
      } else if (h == 0) {  // narrow block: the row owner reads its own elements
        const uint32_t row = live ? (ch.index ? (uint32_t)__ldg(ch.index + row0 + r) : row0 + r) : 0u;
        const float* src = row_ptr(ch.ptr, row, (uint32_t)ch.ld * 4u);
        for (int g8 = 0; g8 < groups; ++g8) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = 8 * g8 + j;
            float x = (live && col < ch.width) ? __ldg(src + col) : 0.f;
            if (ch.relu) x = fmaxf(x, 0.f);
            v[j] = x * rscale;
          }
          split_store8(tm_lane + TM_A_HI + 8 * g8, tm_lane + TM_A_LO + 8 * g8, v);
        }
      }
      tmem_st_wait();
      tc_fence_before_sync();
      team_sync(team);
      TC_PROF(2);
      if (wteam == 0) {  // whole warp, uniform operands (see tc_issue_mmas)
        tc_fence_after_sync();
        if (W64 && ch.staged && (ch.koff & 31) == 0) {  // a 64-wide block starting on a K tile of the weights
          if (team == 0) tc_issue_mmas_k64<0, 0>(wbase, bar_base, p, c == 0, halves0, (uint32_t)(ch.koff >> 5));
          else           tc_issue_mmas_k64<1, 0>(wbase, bar_base, p, c == 0, halves0, (uint32_t)(ch.koff >> 5));
        } else {
          if (team == 0) tc_issue_mmas<0>(wbase, bar_base, p, 0, ch.koff, groups, c == 0);
          else           tc_issue_mmas<1>(wbase, bar_base, p, 0, ch.koff, groups, c == 0);
        }
      }
      TC_PROF(3);
      if (ch.staged) {  // the slot is free: keep the ring full
        consume_done();
        issue_next(t);
      }
      TC_PROF(4);
    }

    // ---------------- hidden layers: accumulator -> bias (+ gathered rows) -> ReLU -> next A operand
    for (int l = 0; l < last; ++l) {
      const float* bias = reinterpret_cast<const float*>(wsm + p.b_off[l]);
      const int nb = W64 ? 4 : (p.npad[l] >> 4), per = (nb + 1) >> 1;
      const int b0 = h * per, b1 = W64 ? b0 + 2 : min(nb, b0 + per);
      uint32_t add_sl[2] = {0u, 0u};         // staged pre-projected blocks: own row in their slots
      const float* add_row[2] = {nullptr, nullptr};
      float4 pre[8];  // the row owner's columns of the first directly read block, fetched under the MMA
#pragma unroll
      for (int q = 0; q < 8; ++q) pre[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      const bool dot_last = p.narrow_last && l == last - 1;  // the last Linear on the CUDA cores
      float part[4] = {0.f, 0.f, 0.f, 0.f};
      int n_staged = 0;
      if (l == 0) {
        bool have_pre = false;
        uint32_t s = cons_slot;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          if (a >= p.n_adds) break;
          if (p.add[a].staged) {
            add_sl[a] = slots + s * TC_SLOT + rsw;
            if (++s == ring) s = 0;
            ++n_staged;
          } else if (live) {
            const uint32_t row = p.add[a].index ? (uint32_t)rcur[1 + a] : row0 + r;
            const float* rowp = row_ptr(p.add[a].ptr, row, (uint32_t)p.add[a].ld * 4u);
            if (!have_pre) {
              have_pre = true;
#pragma unroll
              for (int q = 0; q < 8; ++q) {  // pre[4 bi + q]: 16-byte piece q of this thread's block bi
                const int c4 = 4 * b0 + q;
                if (W64 || (q < 4 * (b1 - b0) && c4 * 4 < p.ntrue[0])) pre[q] = __ldg(reinterpret_cast<const float4*>(rowp) + c4);
              }
            } else {
              add_row[a] = rowp;
            }
          }
        }
      }
      TC_PROF(5);
      if (n_staged > 0) {
        cp_async_wait_pending(issued - consumed - n_staged);
        team_sync(team);
      }
      TC_PROF(7);
      if (l == 0) load_far();
      wait_or_trap(mma_bar, mma_phase);  // every thread, also one without a column block
      mma_phase ^= 1;
      tc_fence_after_sync();
#pragma unroll
      for (int bi = 0; bi < 2; ++bi) {  // at most two 16-column blocks per thread
        const int b = b0 + bi;
        if (!W64 && b >= b1) break;
        uint32_t acc[16];
        tmem_ld16(tm_lane + TM_D + 16 * b, acc);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
        {
          const float4* b4 = reinterpret_cast<const float4*>(bias + 16 * b);
          add16(v, b4[0], b4[1], b4[2], b4[3]);
        }
        if (l == 0) {
          add16(v, pre[bi * 4 + 0], pre[bi * 4 + 1], pre[bi * 4 + 2], pre[bi * 4 + 3]);
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            if (add_sl[a] == 0u && add_row[a] == nullptr) continue;
            float4 x[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int c4 = 4 * b + q;
              x[q] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (W64 || c4 * 4 < p.ntrue[0]) {
                if (add_sl[a]) x[q] = lds128(add_sl[a] + (((uint32_t)c4 << 4) ^ rx));
                else x[q] = __ldg(reinterpret_cast<const float4*>(add_row[a]) + c4);
              }
            }
            add16(v, x[0], x[1], x[2], x[3]);
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
        if (dot_last) {
          const float* wl = reinterpret_cast<const float*>(wsm + p.w_off[last][0]) + 16 * b;
          const int kp = p.kpad[last];
#pragma unroll
          for (int nn = 0; nn < 4; ++nn) {
            if (nn < p.ntrue[last]) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 w4 = *reinterpret_cast<const float4*>(wl + nn * kp + 4 * q);
                part[nn] = fmaf(v[4 * q + 0], w4.x, part[nn]); part[nn] = fmaf(v[4 * q + 1], w4.y, part[nn]);
                part[nn] = fmaf(v[4 * q + 2], w4.z, part[nn]); part[nn] = fmaf(v[4 * q + 3], w4.w, part[nn]);
              }
            }
          }
        } else {
          split_store16(tm_lane + TM_A_HI + 16 * b, tm_lane + TM_A_LO + 16 * b, v);
        }
      }
      if (dot_last) {  // the two column halves of a row meet in the (free) output slot: 16-byte chunks 8 + h
        uint32_t so = cons_slot + (uint32_t)n_staged;
        if (so >= ring) so -= ring;
        sts128(slots + so * TC_SLOT + rsw + (((uint32_t)(8 + h) << 4) ^ rx), make_float4(part[0], part[1], part[2], part[3]));
      }
      tmem_st_wait();
      tc_fence_before_sync();
      team_sync(team);
      TC_PROF(l == 0 ? 8 : 12);
      if (wteam == 0 && !dot_last) {
        tc_fence_after_sync();
        if (W64) {
          const bool hv_next = l + 1 == last && p.npad[last] == 64;  // only the last Linear, see halves0
          if (team == 0) { if (l == 0) tc_issue_mmas_k64<0, 1>(wbase, bar_base, p, true, hv_next); else tc_issue_mmas_k64<0, 2>(wbase, bar_base, p, true, hv_next); }
          else           { if (l == 0) tc_issue_mmas_k64<1, 1>(wbase, bar_base, p, true, hv_next); else tc_issue_mmas_k64<1, 2>(wbase, bar_base, p, true, hv_next); }
        } else {
          if (team == 0) tc_issue_mmas<0>(wbase, bar_base, p, l + 1, 0, p.kpad[l + 1] >> 3, true);
          else           tc_issue_mmas<1>(wbase, bar_base, p, l + 1, 0, p.kpad[l + 1] >> 3, true);
        }
      }
      TC_PROF(9);
      for (int a = 0; a < n_staged; ++a) {  // the slots of the staged pre-projected blocks are free
        consume_done();
        issue_next(t);
      }
      TC_PROF(10);
    }

    // ---------------- output: accumulator -> bias -> activation -> staged tile
    const uint32_t osl = slots + cons_slot * TC_SLOT;  // this item's slot was released `ring` items ago
    if (p.narrow_last) {
      if (h == 0) {
        const float* bias = reinterpret_cast<const float*>(wsm + p.b_off[last]);
        const float4 a = lds128(osl + rsw + ((8u << 4) ^ rx)), b = lds128(osl + rsw + ((9u << 4) ^ rx));
        float v[4] = {a.x + b.x + bias[0], a.y + b.y + bias[1], a.z + b.z + bias[2], a.w + b.w + bias[3]};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (p.final_act == GTB_ACT_RELU) v[j] = fmaxf(v[j], 0.f);
          else if (p.final_act == GTB_ACT_SIGMOID_AFFINE) v[j] = p.act_eps + (1.f - 2.f * p.act_eps) * (1.f / (1.f + expf(-v[j])));
        }
        sts128(osl + rsw + rx, make_float4(v[0], v[1], v[2], v[3]));  // chunk 0 of the own row
      }
    }
    TC_PROF(13);
    if (!p.narrow_last) {
      const float* bias = reinterpret_cast<const float*>(wsm + p.b_off[last]);
      const int nb = p.npad[last] >> 4, per = (nb + 1) >> 1;
      const int b0 = h * per, b1 = min(nb, b0 + per);
      // the last accumulator arrives in two halves when it was issued that way (see halves0 / hv_next)
      const bool hv = W64 && p.npad[last] == 64 && (last > 0 || halves0);
      if (last == 0) load_far();
      wait_or_trap(mma_bar, mma_phase);
      mma_phase ^= 1;
      tc_fence_after_sync();
      for (int bi = 0; bi < 2; ++bi) {
        const int b = hv ? (bi == 0 ? h : 2 + h) : b0 + bi;
        if (!hv && b >= b1) break;
        if (hv && bi == 1) {
          wait_or_trap(mma_bar1, mma_phase1);
          mma_phase1 ^= 1;
          tc_fence_after_sync();
        }
        uint32_t acc[16];
        tmem_ld16(tm_lane + TM_D + 16 * b, acc);
        tmem_ld_wait();
        float vv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) vv[j] = __uint_as_float(acc[j]);
        {
          const float4* b4 = reinterpret_cast<const float4*>(bias + 16 * b);
          add16(vv, b4[0], b4[1], b4[2], b4[3]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float x = vv[4 * q + j];
            if (p.final_act == GTB_ACT_RELU) x = fmaxf(x, 0.f);
            else if (p.final_act == GTB_ACT_SIGMOID_AFFINE) x = p.act_eps + (1.f - 2.f * p.act_eps) * (1.f / (1.f + expf(-x)));
            v[j] = x;
          }
          sts128(osl + rsw + (((uint32_t)(4 * b + q) << 4) ^ rx), make_float4(v[0], v[1], v[2], v[3]));
        }
      }
    }
    tc_fence_before_sync();
    team_sync(team);
    TC_PROF(14);

    // ---------------- residual / scale, coalesced (scattered) row stores
    const int N = p.ntrue[last];
    const bool want_aggr = p.aggr != nullptr;
    const float oscale = p.out_scale ? __ldg(p.out_scale) : 1.f;
    const bool touch = p.res != nullptr || p.res_b != 1.f || p.out_scale != nullptr;
    // a gate (backward launches: the ReLU mask of the recomputed activation) rides on the 64-wide vector path;
    // other widths take the plain scalar path below
    const bool gate_vec = p.gate == nullptr || (N == 64 && p.out_mode != 2 && (p.gate_ld & 3) == 0 &&
                                                (reinterpret_cast<uintptr_t>(p.gate) & 15) == 0);
    const bool vec = gate_vec && (N & 3) == 0 &&
                     (p.out == nullptr || ((p.out_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)) &&
                     (p.res == nullptr || ((p.res_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.res) & 15) == 0));
    if (p.out != nullptr || touch || p.gate != nullptr) {
      if (vec && p.out_mode != 2 && N == 64) {  // 16 pieces per row: thread tt stores piece (tt & 15) of rows (tt >> 4) + 16 j
        const int rb = tt >> 4, c = tt & 15;
        const uint32_t sp0 = osl + slot_off(rb, c);
        const uint32_t old4 = (uint32_t)p.out_ld * 4u, rld4 = (uint32_t)p.res_ld * 4u;
        const float* res0 = p.res ? row_ptr(p.res + (c << 2), row0 + rb, rld4) : nullptr;
        const f32x2 resa2 = pack2(p.res_a, p.res_a), resb2 = pack2(p.res_b, p.res_b), osc2 = pack2(oscale, oscale);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rr = rb + 16 * j;
          uint32_t orow = row0 + rr;
          if (p.out_mode == 1) orow = (uint32_t)__shfl_sync(0xffffffffu, ocur, 2 * j + (lane >> 4));
          if (rr < rows_here) {
            float4 v = lds128(sp0 + j * 4096);
            if (touch) {  // (res_a res + res_b v) oscale, same operation order as the scalar paths
              f32x2 lo2 = mul2(pack2(v.x, v.y), resb2), hi2 = mul2(pack2(v.z, v.w), resb2);
              if (p.res) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(row_ptr(res0, 16 * j, rld4)));
                lo2 = fma2(resa2, pack2(q.x, q.y), lo2);
                hi2 = fma2(resa2, pack2(q.z, q.w), hi2);
              }
              if (p.out_scale) {
                lo2 = mul2(lo2, osc2);
                hi2 = mul2(hi2, osc2);
              }
              unpack2(lo2, v.x, v.y);
              unpack2(hi2, v.z, v.w);
              if (want_aggr) sts128(sp0 + j * 4096, v);
            }
            if (p.gate) {  // out = gate > 0 ? value : 0 (never combined with an aggregate)
              const float4 gq = __ldg(reinterpret_cast<const float4*>(row_ptr(p.gate + (c << 2), row0 + rr, (uint32_t)p.gate_ld * 4u)));
              v.x = gq.x > 0.f ? v.x : 0.f; v.y = gq.y > 0.f ? v.y : 0.f;
              v.z = gq.z > 0.f ? v.z : 0.f; v.w = gq.w > 0.f ? v.w : 0.f;
            }
            if (p.out) *reinterpret_cast<float4*>(const_cast<float*>(row_ptr(p.out + (c << 2), orow, old4))) = v;
          }
        }
      } else if (vec && p.out_mode != 2) {
        const int c4n = N >> 2;  // a power of two here (out_mode 0 / 1)
        const int sh = __ffs(c4n) - 1;
        const int total = rows_here * c4n;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int i = tt + j * TC_TEAM;
          const int rr = i >> sh, c = i & (c4n - 1);
          uint32_t orow = row0 + rr;
          if (p.out_mode == 1) orow = (uint32_t)__shfl_sync(0xffffffffu, ocur, (j * (32 >> sh) + (lane >> sh)) & 31);
          if (i < total) {
            const uint32_t sp = osl + slot_off(rr, c);
            float4 v = lds128(sp);
            if (touch) {
              v.x *= p.res_b; v.y *= p.res_b; v.z *= p.res_b; v.w *= p.res_b;
              if (p.res) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(row_ptr(p.res, row0 + rr, (uint32_t)p.res_ld * 4u)) + c);
                v.x = fmaf(p.res_a, q.x, v.x); v.y = fmaf(p.res_a, q.y, v.y);
                v.z = fmaf(p.res_a, q.z, v.z); v.w = fmaf(p.res_a, q.w, v.w);
              }
              v.x *= oscale; v.y *= oscale; v.z *= oscale; v.w *= oscale;
              if (want_aggr) sts128(sp, v);
            }
            if (p.out) *(reinterpret_cast<float4*>(const_cast<float*>(row_ptr(p.out, orow, (uint32_t)p.out_ld * 4u))) + c) = v;
          }
        }
      } else if (N == 1 && p.out_mode != 2 && p.gate == nullptr) {  // one value per row (edge weights): thread = row
        uint32_t orow = row0 + tt;
        if (p.out_mode == 1) orow = (uint32_t)ocur;  // c4n = 1: every lane holds its own row
        if (tt < rows_here) {
          const uint32_t sp = osl + slot_off(tt, 0);
          float v = lds32(sp);
          if (touch) {
            v *= p.res_b;
            if (p.res) v = fmaf(p.res_a, __ldg(row_ptr(p.res, row0 + tt, (uint32_t)p.res_ld * 4u)), v);
            v *= oscale;
            if (want_aggr) sts32(sp, v);
          }
          if (p.out) *const_cast<float*>(row_ptr(p.out, orow, (uint32_t)p.out_ld * 4u)) = v;
        }
      } else {
        for (int i = tt; i < rows_here * N; i += TC_TEAM) {
          const int rr = i / N, n = i - rr * N;
          const uint32_t sp = osl + slot_off(rr, n >> 2) + (n & 3) * 4;
          float v = lds32(sp);
          if (touch) {
            v *= p.res_b;
            if (p.res) v = fmaf(p.res_a, __ldg(row_ptr(p.res, row0 + rr, (uint32_t)p.res_ld * 4u) + n), v);
            v *= oscale;
            if (want_aggr) sts32(sp, v);
          }
          if (p.gate && !(__ldg(row_ptr(p.gate, row0 + rr, (uint32_t)p.gate_ld * 4u) + n) > 0.f)) v = 0.f;
          if (p.out) {
            const uint32_t orow = p.out_index ? (uint32_t)__ldg(p.out_index + row0 + rr) : row0 + rr;
            const_cast<float*>(row_ptr(p.out, orow, (uint32_t)p.out_ld * 4u))[n] = v;
          }
        }
      }
    }
    TC_PROF(15);

    // ---------------- in-tile segmented sum by destination (rows are destination-sorted): thread =
    // (column, 32-row quarter); a warp shares its quarter, so the run boundaries are warp-uniform
    // branches.  One fire-and-forget reduction per (run, column) onto the zero-filled aggregate.
    const bool aggr_vec = want_aggr && (N & 3) == 0 && (p.aggr_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.aggr) & 15) == 0;
    if (aggr_vec) {
      // thread = (4 adjacent columns, group of 8 rows): 8 x 16-byte shared loads, one vector
      // reduction (red.global.add.v4.f32) per run of equal destinations inside the group
      if (touch) team_sync(team);
      const int c4 = tt & 15, rg0 = (tt >> 4) * 8;
      if (c4 * 4 < N && rg0 < rows_here) {
        const uint32_t al4 = (uint32_t)p.aggr_ld * 4u;
        int sg[8];
        {
          const uint32_t sa = segs + (uint32_t)rg0 * 4u;
          int4 s0, s1;
          asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(s0.x), "=r"(s0.y), "=r"(s0.z), "=r"(s0.w) : "r"(sa) : "memory");
          asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(s1.x), "=r"(s1.y), "=r"(s1.z), "=r"(s1.w) : "r"(sa + 16) : "memory");
          sg[0] = s0.x; sg[1] = s0.y; sg[2] = s0.z; sg[3] = s0.w; sg[4] = s1.x; sg[5] = s1.y; sg[6] = s1.z; sg[7] = s1.w;
        }
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)  // (rg0 + i) & 7 == i: rg0 is a multiple of 8
          v[i] = lds128(osl + (uint32_t)(rg0 + i) * 256u + (((uint32_t)c4 << 4) ^ ((uint32_t)i << 4)));
        int cur = sg[0];
        f32x2 s01 = pack2(0.f, 0.f), s23 = s01;
        auto flush = [&](int seg) {
          float4 sum;
          unpack2(s01, sum.x, sum.y);
          unpack2(s23, sum.z, sum.w);
          red_add_v4(row_ptr(p.aggr + 4 * c4, (uint32_t)seg, al4), sum);
        };
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (sg[i] < 0) break;  // rows past the end of the last tile
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
    } else if (want_aggr) {
      if (touch) team_sync(team);
      const int c = tt & 63, r0 = (tt >> 6) * 32;
      // run starts of this warp's 32 rows as a warp-uniform bit mask; all 32 values of the column
      // are loaded before the (branchy, but uniform) summation so that their latencies overlap
      const int myrow = r0 + lane;
      const int myseg = myrow < rows_here ? lds_i32(segs + myrow * 4) : -1;
      const int prevseg = __shfl_up_sync(0xffffffffu, myseg, 1);
      const uint32_t starts = __ballot_sync(0xffffffffu, myrow < rows_here && (lane == 0 || myseg != prevseg));
      if (r0 < rows_here) {  // warp-uniform; `act` is not (N < 64): the shuffles below stay outside of it
        const bool act = c < N;
        const uint32_t al4 = (uint32_t)p.aggr_ld * 4u;
        const uint32_t cx = (uint32_t)(c >> 2) << 4, cw = (uint32_t)(c & 3) * 4u;
        const uint32_t col0 = osl + (uint32_t)r0 * 256u + cw;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i)  // (r0 + i) & 7 == i & 7: r0 is a multiple of 32
          v[i] = (act && r0 + i < rows_here) ? lds32(col0 + (uint32_t)i * 256u + (cx ^ ((uint32_t)(i & 7) << 4))) : 0.f;
        int cur = __shfl_sync(0xffffffffu, myseg, 0);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i > 0 && ((starts >> i) & 1u)) {  // warp-uniform
            if (act) atomicAdd(const_cast<float*>(row_ptr(p.aggr + c, (uint32_t)cur, al4)), sum);
            cur = __shfl_sync(0xffffffffu, myseg, i);
            sum = 0.f;
          }
          sum += v[i];
        }
        if (act) atomicAdd(const_cast<float*>(row_ptr(p.aggr + c, (uint32_t)cur, al4)), sum);
      }
    }
    TC_PROF(16);
    team_sync(team);
    consume_done();
    issue_next(t);  // the output slot is free again
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ccur[q] = cnext[q];
      cnext[q] = cfar[q];
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      rcur[q] = rnext[q];
      rnext[q] = rfar[q];
    }
    ocur = onext;
    onext = ofar;
    TC_PROF(17);
    if (PROF && prof_on) g_tc_prof[31] += 1;
  }

  cp_async_wait_pending(0);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, tmem_cols);
}

// ------------------------------------------------------------------------------ host side
bool tc_supported(int n_layers, const int32_t* dims, int n_chunks, const int32_t* chunk_w) {
  TcLayout L;
  return tc_layout(n_layers, dims, n_chunks, chunk_w, &L);
}

int fused_mlp_tc(const gtb_mlp_desc_t& d, cudaStream_t st) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  int32_t chunk_w[GTB_MAX_SRCS];
  int n_stream = 0;
  for (int s = 0; s < d.n_srcs; ++s)
    if (!(d.srcs[s].flags & GTB_SRC_PROJECTED)) chunk_w[n_stream++] = d.srcs[s].width;
  TcLayout L;
  GTB_REQUIRE(n_stream >= 1 && tc_layout(d.n_layers, d.dims, n_stream, chunk_w, &L), GTB_ERR_UNSUPPORTED_DIM,
              "gtb_fused_mlp_f32: these widths are not supported by the tcgen05 path");
  p.n_rows = d.n_rows;
  p.n_tiles = (int32_t)((d.n_rows + TC_TM - 1) / TC_TM);
  p.n_layers = d.n_layers;
  int n_items = 0, c = 0;
  for (int s = 0; s < d.n_srcs; ++s) {
    const gtb_src_t& src = d.srcs[s];
    if (src.flags & GTB_SRC_PROJECTED) continue;
    const int ci = c++;
    const bool staged = (src.width & 3) == 0 && (src.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src.ptr) & 15) == 0;
    GTB_REQUIRE(staged || src.width <= 16, GTB_ERR_UNSUPPORTED_DIM,
                "gtb_fused_mlp_f32 (tcgen05): source block %d (width %d, ld %d) is neither 16-byte aligned nor narrow", s,
                src.width, src.ld);
    // blocks wider than 64 columns go through the A buffer in 64-column pieces
    for (int off = 0; off < src.width; off += 64) {
      GTB_REQUIRE(p.n_chunks < TC_MAXCH, GTB_ERR_UNSUPPORTED_DIM, "gtb_fused_mlp_f32 (tcgen05): too many streamed blocks");
      TcChunk& ch = p.ch[p.n_chunks];
      ch.ptr = src.ptr + off;
      ch.index = src.index;
      ch.ld = src.ld;
      ch.width = src.width - off < 64 ? src.width - off : 64;
      ch.kpad = round_up(ch.width, 8);
      ch.koff = L.chunk_koff[ci] + off;
      ch.relu = src.relu;
      ch.staged = staged;
      if (staged) p.items[n_items++] = (int8_t)p.n_chunks;
      ++p.n_chunks;
    }
  }
  for (int s = 0; s < d.n_srcs; ++s) {
    const gtb_src_t& src = d.srcs[s];
    if (!(src.flags & GTB_SRC_PROJECTED)) continue;
    GTB_REQUIRE(d.n_layers >= 2 && p.n_adds < 2, GTB_ERR_UNSUPPORTED_DIM,
                "gtb_fused_mlp_f32 (tcgen05): at most two pre-projected blocks, and only in front of a hidden layer");
    GTB_REQUIRE(src.width == d.dims[1] && (src.width & 3) == 0 && (src.ld & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(src.ptr) & 15) == 0 && !src.relu,
                GTB_ERR_BAD_ARG, "gtb_fused_mlp_f32 (tcgen05): pre-projected block %d must be [*, %d] fp32, 16-byte aligned", s,
                d.dims[1]);
    TcAdd& a = p.add[p.n_adds];
    a.ptr = src.ptr;
    a.index = src.index;
    a.ld = src.ld;
    a.staged = !((src.flags & GTB_SRC_SORTED) || src.index == nullptr);
    if (a.staged) p.items[n_items++] = (int8_t)(64 + p.n_adds);
    ++p.n_adds;
  }
  p.items[n_items++] = -1;  // the output tile
  p.ipt = n_items;
  // row-index registers of the gather copies: one per distinct (index array, pieces per row)
  for (int k = 0; k < n_items; ++k) {
    const int kind = p.items[k];
    p.item_ireg[k] = TC_IDX_IDENTITY;
    if (kind < 0) continue;
    const int32_t* index = kind < 64 ? p.ch[kind].index : p.add[kind - 64].index;
    const int c4n = (kind < 64 ? p.ch[kind].width : d.dims[1]) >> 2;
    if (index == nullptr) continue;
    p.item_ireg[k] = TC_IDX_GLOBAL;
    if ((c4n & (c4n - 1)) != 0) continue;
    int q = 0;
    while (q < p.n_iregs && !(p.ireg_ptr[q] == index && p.ireg_c4n[q] == c4n)) ++q;
    if (q == p.n_iregs) {
      if (q == 4) continue;
      p.ireg_ptr[q] = index;
      p.ireg_c4n[q] = c4n;
      ++p.n_iregs;
    }
    p.item_ireg[k] = (int8_t)q;
  }
  {
    const int n_out = d.dims[d.n_layers];
    const int c4n = (n_out & 3) == 0 ? n_out >> 2 : (n_out == 1 ? 1 : 0);
    const bool regular = c4n > 0 && (c4n & (c4n - 1)) == 0;
    p.out_c4n = regular ? c4n : 1;
    p.out_mode = regular ? (d.out_index ? 1 : 0) : 2;
  }
  p.prof = g_prof_enabled;
  for (int l = 0; l < d.n_layers; ++l) {
    p.kpad[l] = L.kpad[l];
    p.npad[l] = L.npad[l];
    p.ntrue[l] = L.ntrue[l];
    p.w_off[l][0] = L.w_off[l][0];
    p.w_off[l][1] = L.w_off[l][1];
    p.b_off[l] = L.b_off[l];
  }
  p.w_bytes = L.total_bytes;
  p.narrow_last = L.narrow_last;
  p.packed = static_cast<const unsigned char*>(d.packed);
  GTB_REQUIRE((reinterpret_cast<uintptr_t>(d.packed) & 15) == 0, GTB_ERR_BAD_ARG, "gtb_fused_mlp_f32: packed weights must be 16-byte aligned");
  p.final_act = d.final_act;
  p.act_eps = d.act_eps;
  p.res_a = d.res_a;
  p.res_b = d.res_b;
  p.res = d.res;
  p.res_ld = d.res_ld;
  p.row_scale = d.row_scale;
  p.out_scale = d.out_scale;
  p.gate = d.gate;
  p.gate_ld = d.gate_ld;
  p.out = d.out;
  p.out_index = d.out_index;
  p.out_ld = d.out_ld;
  p.aggr = d.aggr;
  p.aggr_ld = d.aggr_ld;
  p.seg_id = d.seg_id;
  p.rowptr = d.rowptr;
  const size_t fixed = ((size_t)L.total_bytes + 15) / 16 * 16 + TC_MISC;
  const int n_slots = (int)((TC_SMEM_MAX - fixed) / TC_SLOT);
  GTB_REQUIRE(n_slots >= 1, GTB_ERR_UNSUPPORTED_DIM, "gtb_fused_mlp_f32 (tcgen05): weights do not fit in shared memory");
  // two teams whenever there are two slots: with one slot each, a team's gather latency is exposed
  // but overlaps the other team's conversion / epilogue work (measured better than one team with a
  // two-slot ring on the W head); GTB_TC_TEAMS=1 forces one team (experiments)
  static const int force_teams = getenv("GTB_TC_TEAMS") ? atoi(getenv("GTB_TC_TEAMS")) : 0;
  p.n_teams = (n_slots >= 2 && p.n_tiles > kNumSMs) ? 2 : 1;
  if (force_teams == 1) p.n_teams = 1;
  p.ring = n_slots / p.n_teams;
  if (p.ring > 4) p.ring = 4;
  if (p.ring > p.ipt) p.ring = p.ipt;  // at most one tile of look-ahead: the index registers hold one tile
  bool staged_add = false;
  for (int a2 = 0; a2 < p.n_adds; ++a2) staged_add = staged_add || p.add[a2].staged;
  if (p.narrow_last && d.n_layers == 2 && staged_add && p.ring < 2) {
    // the last hidden epilogue then writes its partial dot products into the output slot while it
    // still reads a staged block: they must be different slots
    GTB_REQUIRE(n_slots >= 2, GTB_ERR_UNSUPPORTED_DIM, "gtb_fused_mlp_f32 (tcgen05): not enough shared memory for this shape");
    p.n_teams = 1;
    p.ring = n_slots < p.ipt ? n_slots : p.ipt;
    if (p.ring > 4) p.ring = 4;
  }
  if (d.n_rows == 0) return GTB_OK;
  const size_t smem = fixed + (size_t)p.ring * p.n_teams * TC_SLOT;
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fused_mlp_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fused_mlp_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fused_mlp_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fused_mlp_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(fused_mlp_tc)");
    configured = true;
  }
  bool w64 = true;  // the straight-line variant: everything in front of the last Linear is 64 wide
  for (int c2 = 0; c2 < p.n_chunks; ++c2) w64 = w64 && (!p.ch[c2].staged || p.ch[c2].width == 64);
  for (int l = 0; l + 1 < d.n_layers; ++l) w64 = w64 && d.dims[l + 1] == 64;
  const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
  const int nthr = TC_TEAM * p.n_teams;
  if (p.prof) {
    if (w64) fused_mlp_tc_kernel<true, true><<<grid, nthr, smem, st>>>(p);
    else     fused_mlp_tc_kernel<false, true><<<grid, nthr, smem, st>>>(p);
  } else {
    if (w64) fused_mlp_tc_kernel<true, false><<<grid, nthr, smem, st>>>(p);
    else     fused_mlp_tc_kernel<false, false><<<grid, nthr, smem, st>>>(p);
  }
  GTB_CHECK_LAUNCH("fused_mlp_tc_kernel");
  return GTB_OK;
}

int tc_profile(int enable, long long* out32) {
  g_prof_enabled = enable;
  if (out32 == nullptr) {
    long long zero[32] = {0};
    return check_cuda(cudaMemcpyToSymbol(g_tc_prof, zero, sizeof(zero)), "cudaMemcpyToSymbol(g_tc_prof)");
  }
  return check_cuda(cudaMemcpyFromSymbol(out32, g_tc_prof, 32 * sizeof(long long)), "cudaMemcpyFromSymbol(g_tc_prof)");
}

int tc_timeout_flag(int* out) {
  return check_cuda(cudaMemcpyFromSymbol(out, g_tc_timeout, sizeof(int)), "cudaMemcpyFromSymbol(g_tc_timeout)");
}

}  // namespace gtb
