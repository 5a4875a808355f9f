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
#include "common.cuh"
#include "tc_common.cuh"

namespace gtb {

using namespace tc;

constexpr int TC_TM = 128;                 // rows per tile = UMMA M
constexpr int TC_NT = 256;                 // threads: (row = tid & 127, column half = tid >> 7)
constexpr int TC_SLOT = TC_TM * 256;       // one staging slot: [128 rows][64 fp32], 16-byte chunks ^ (row & 7)
constexpr int TC_MAXCH = 8;                // streamed sub-blocks (<= 64 columns each) of the first Linear
constexpr int TC_MAXITEMS = TC_MAXCH + 3;
constexpr int TC_SMEM_MAX = 232448;        // 227 KB opt-in shared memory per CTA
constexpr int TC_HEAD = 1040;              // row indices (2 x 128 int32), MMA barrier, TMEM slot
constexpr int TC_MISC = TC_HEAD + 1023;    // + slack to align the weight tiles to 1024 bytes
constexpr uint32_t TM_A_HI = 0, TM_A_LO = 64, TM_D = 128;  // TMEM column map of one tile context

__device__ int g_tc_timeout = 0;

// packed weights of one MLP: per layer the tf32 hi and lo parts as K-major [npad][32] fp32 tiles
// (128-byte swizzle), then one 64-float bias row per layer.  Layer 0 keeps each streamed column
// block padded to a multiple of 8 columns.
struct TcLayout {
  int n_layers, n_chunks;
  int chunk_w[GTB_MAX_SRCS], chunk_koff[GTB_MAX_SRCS], chunk_kpad[GTB_MAX_SRCS];
  int kpad[GTB_MAX_LAYERS], npad[GTB_MAX_LAYERS], ntrue[GTB_MAX_LAYERS], ktiles[GTB_MAX_LAYERS];
  uint32_t w_off[GTB_MAX_LAYERS][2], b_off[GTB_MAX_LAYERS], total_bytes;
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
  int32_t n_tiles, n_chunks, n_adds, n_layers, ring, ipt;
  int8_t items[TC_MAXITEMS + 1];  // per tile, in consumption order: chunk c -> c, add a -> 64 + a, output tile -> -1
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
  uint32_t off = 0;
  for (int l = 0; l < n_layers; ++l) {
    const int n = dims[l + 1];
    if (n < 1 || n > 64) return false;
    L->ntrue[l] = n;
    L->npad[l] = round_up(n, 16);
    L->kpad[l] = (l == 0) ? k0 : L->npad[l - 1];
    L->ktiles[l] = (L->kpad[l] + 31) / 32;
    const uint32_t bytes = (uint32_t)L->ktiles[l] * L->npad[l] * 128u;  // multiple of 2048
    L->w_off[l][0] = off;
    off += bytes;
    L->w_off[l][1] = off;
    off += bytes;
  }
  for (int l = 0; l < n_layers; ++l) {
    L->b_off[l] = off;
    off += 64 * 4;
  }
  L->total_bytes = off;
  return (size_t)off + TC_SLOT + TC_MISC <= (size_t)TC_SMEM_MAX;
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
__device__ __forceinline__ uint32_t slot_off(int r, int c4) {  // 16-byte chunk c4 of row r
  return (uint32_t)(r * 256 + ((c4 ^ (r & 7)) << 4));
}

__device__ __forceinline__ void wait_or_trap(uint64_t* bar, uint32_t parity) {
  if (!mbar_wait(bar, parity, 20000000u)) {
    atomicExch(&g_tc_timeout, 1);
    __trap();  // a wrong descriptor must fail loudly, never hang the GPU or return garbage
  }
}

// gather one staged item (a streamed column block or a pre-projected row block) of sequence
// number g into its ring slot; every thread commits exactly one cp.async group per call
__device__ __forceinline__ void tc_issue_item(const TcParams& p, int g, unsigned char* slots, int tid) {
  const int t = g / p.ipt, k = g - t * p.ipt;
  const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
  const int kind = p.items[k];
  if (tile < p.n_tiles && kind >= 0) {
    const float* ptr;
    const int32_t* index;
    int ld, width;
    if (kind < 64) {
      ptr = p.ch[kind].ptr; index = p.ch[kind].index; ld = p.ch[kind].ld; width = p.ch[kind].width;
    } else {
      ptr = p.add[kind - 64].ptr; index = p.add[kind - 64].index; ld = p.add[kind - 64].ld; width = p.ntrue[0];
    }
    const int64_t row0 = tile * TC_TM;
    const int rows_here = (int)min((int64_t)TC_TM, p.n_rows - row0);
    const uint32_t sbase = smem_u32(slots + (size_t)(g % p.ring) * TC_SLOT);
    const int c4n = width >> 2;
    const int total = rows_here * c4n;
    for (int i = tid; i < total; i += TC_NT) {
      const int r = i / c4n, c = i - r * c4n;
      const int64_t row = index ? (int64_t)__ldg(index + row0 + r) : row0 + r;
      cp_async16(sbase + slot_off(r, c), ptr + (size_t)row * ld + (c << 2));
    }
  }
  cp_async_commit();
}

__device__ __forceinline__ void tc_issue_mmas(uint32_t tm, uint32_t wbase, const TcParams& p, int l, int koff,
                                              int ksteps, bool first) {
  const uint32_t idesc = make_idesc_tf32(TC_TM, p.npad[l]);
  const uint32_t tile_bytes = (uint32_t)p.npad[l] * 128u;
  bool acc = !first;
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {  // small terms first: lo*hi, hi*lo, hi*hi
    const uint32_t a_col = (pass == 0) ? TM_A_LO : TM_A_HI;
    const uint32_t b_base = wbase + p.w_off[l][pass == 1 ? 1 : 0];
#pragma unroll 1
    for (int ks = 0; ks < ksteps; ++ks) {
      const int kg = koff + 8 * ks;
      const uint64_t bd = make_smem_desc_sw128(b_base + (uint32_t)(kg >> 5) * tile_bytes + (uint32_t)((kg & 31) >> 3) * 32u);
      mma_tf32_ts(tm + TM_D, tm + a_col + 8 * ks, bd, idesc, acc);
      acc = true;
    }
  }
}

__device__ __forceinline__ void split_store8(uint32_t taddr_hi, uint32_t taddr_lo, const float (&v)[8]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float h, l;
    split_tf32(v[j], h, l);
    hi[j] = __float_as_uint(h);
    lo[j] = __float_as_uint(l);
  }
  tmem_st8(taddr_hi, hi);
  tmem_st8(taddr_lo, lo);
}

__global__ void __launch_bounds__(TC_NT, 1) fused_mlp_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  int32_t* orow = reinterpret_cast<int32_t*>(smem_raw);
  int32_t* segs = orow + TC_TM;
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(segs + TC_TM);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
  unsigned char* wsm = smem_raw + TC_HEAD;                     // packed weights + biases (1024-aligned tiles)
  wsm += (1024u - (smem_u32(wsm) & 1023u)) & 1023u;
  unsigned char* slots = wsm + ((p.w_bytes + 15u) & ~15u);     // ring of staging slots

  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & (TC_TM - 1), h = tid >> 7;

  // ---- prologue: first items in flight, weights into shared memory, barrier + TMEM set-up
  int issued = 0;
  for (; issued < p.ring; ++issued) tc_issue_item(p, issued, slots, tid);
  {
    const float4* g4 = reinterpret_cast<const float4*>(p.packed);
    float4* s4 = reinterpret_cast<float4*>(wsm);
    for (int i = tid; i < (int)(p.w_bytes >> 4); i += TC_NT) s4[i] = __ldg(g4 + i);
  }
  if (tid == 0) {
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  fence_proxy_async_smem();  // weights were written through the generic proxy, the MMA reads them through the async proxy
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t tm_lane = tm + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t wbase = smem_u32(wsm);
  uint32_t mma_phase = 0;
  const int last = p.n_layers - 1;

  int t = 0;
  for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t) {
    const int64_t row0 = tile * TC_TM;
    const int rows_here = (int)min((int64_t)TC_TM, p.n_rows - row0);
    const bool live = r < rows_here;
    int g = t * p.ipt;  // sequence number of this tile's next staged item
    if (tid < TC_TM) {
      int o = 0, sg = -1;
      if (live) {
        o = p.out_index ? __ldg(p.out_index + row0 + r) : (int)(row0 + r);
        if (p.seg_id) sg = __ldg(p.seg_id + row0 + r);
      }
      orow[tid] = o;
      segs[tid] = sg;
    }
    const float rscale = (p.row_scale && live) ? __ldg(p.row_scale + row0 + r) : 1.f;

    // ---------------- first Linear: one streamed block at a time through the TMEM A buffer
    for (int c = 0; c < p.n_chunks; ++c) {
      const TcChunk& ch = p.ch[c];
      const int groups = ch.kpad >> 3;
      if (ch.staged) {
        cp_async_wait_pending(issued - g - 1);
        __syncthreads();
      }
      if (c > 0) {  // the previous block's MMAs still read the A buffer
        wait_or_trap(mma_bar, mma_phase);
        mma_phase ^= 1;
        tc_fence_after_sync();
      }
      if (ch.staged) {
        const unsigned char* sl = slots + (size_t)(g % p.ring) * TC_SLOT;
        for (int g8 = h; g8 < groups; g8 += 2) {
          float v[8];
          const int c4 = 2 * g8;
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
          if (c4 * 4 < ch.width) a = *reinterpret_cast<const float4*>(sl + slot_off(r, c4));
          if ((c4 + 1) * 4 < ch.width) b = *reinterpret_cast<const float4*>(sl + slot_off(r, c4 + 1));
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (ch.relu) v[j] = fmaxf(v[j], 0.f);
            v[j] *= rscale;
          }
          split_store8(tm_lane + TM_A_HI + 8 * g8, tm_lane + TM_A_LO + 8 * g8, v);
        }
      } else if (h == 0) {  // narrow block: the row owner reads its own elements
        const int64_t row = live ? (ch.index ? (int64_t)__ldg(ch.index + row0 + r) : row0 + r) : 0;
        const float* src = ch.ptr + (size_t)row * ch.ld;
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
      __syncthreads();
      if (ch.staged) {  // the slot is free: keep the ring full
        tc_issue_item(p, issued, slots, tid);
        ++issued;
        ++g;
      }
      if (tid == 0) {
        tc_fence_after_sync();
        tc_issue_mmas(tm, wbase, p, 0, ch.koff, groups, c == 0);
        mma_commit(mma_bar);
      }
    }

    // ---------------- hidden layers: accumulator -> bias (+ gathered rows) -> ReLU -> next A operand
    for (int l = 0; l < last; ++l) {
      wait_or_trap(mma_bar, mma_phase);
      mma_phase ^= 1;
      tc_fence_after_sync();
      const float* bias = reinterpret_cast<const float*>(wsm + p.b_off[l]);
      const int groups = p.npad[l] >> 3, half = groups >> 1;
      const unsigned char* add_sl[2] = {nullptr, nullptr};
      const float* add_row[2] = {nullptr, nullptr};
      if (l == 0) {
        int gg = g;
        for (int a = 0; a < p.n_adds; ++a) {
          if (p.add[a].staged) {
            add_sl[a] = slots + (size_t)(gg % p.ring) * TC_SLOT;
            ++gg;
          } else if (live) {
            const int64_t row = p.add[a].index ? (int64_t)__ldg(p.add[a].index + row0 + r) : row0 + r;
            add_row[a] = p.add[a].ptr + (size_t)row * p.add[a].ld;
          }
        }
        if (gg > g) {
          cp_async_wait_pending(issued - gg);
          __syncthreads();
        }
      }
      for (int g8 = h * half; g8 < (h + 1) * half; ++g8) {
        uint32_t acc[8];
        tmem_ld8(tm_lane + TM_D + 8 * g8, acc);
        tmem_ld_wait();
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[j]) + bias[8 * g8 + j];
        if (l == 0) {
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
            const int c4 = 2 * g8;
            if (add_sl[a]) {
              if (c4 * 4 < p.ntrue[0]) x = *reinterpret_cast<const float4*>(add_sl[a] + slot_off(r, c4));
              if ((c4 + 1) * 4 < p.ntrue[0]) y = *reinterpret_cast<const float4*>(add_sl[a] + slot_off(r, c4 + 1));
            } else if (add_row[a]) {
              if (c4 * 4 < p.ntrue[0]) x = __ldg(reinterpret_cast<const float4*>(add_row[a]) + c4);
              if ((c4 + 1) * 4 < p.ntrue[0]) y = __ldg(reinterpret_cast<const float4*>(add_row[a]) + c4 + 1);
            }
            v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w; v[4] += y.x; v[5] += y.y; v[6] += y.z; v[7] += y.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        split_store8(tm_lane + TM_A_HI + 8 * g8, tm_lane + TM_A_LO + 8 * g8, v);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncthreads();
      if (l == 0) {
        for (int a = 0; a < p.n_adds; ++a)
          if (p.add[a].staged) {
            tc_issue_item(p, issued, slots, tid);
            ++issued;
            ++g;
          }
      }
      if (tid == 0) {
        tc_fence_after_sync();
        tc_issue_mmas(tm, wbase, p, l + 1, 0, p.kpad[l + 1] >> 3, true);
        mma_commit(mma_bar);
      }
    }

    // ---------------- output: accumulator -> bias -> activation -> staged tile
    wait_or_trap(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after_sync();
    unsigned char* osl = slots + (size_t)(g % p.ring) * TC_SLOT;  // this item's slot was released R items ago
    {
      const float* bias = reinterpret_cast<const float*>(wsm + p.b_off[last]);
      const int groups = p.npad[last] >> 3, half = groups >> 1;
      for (int g8 = h * half; g8 < (h + 1) * half; ++g8) {
        uint32_t acc[8];
        tmem_ld8(tm_lane + TM_D + 8 * g8, acc);
        tmem_ld_wait();
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float x = __uint_as_float(acc[j]) + bias[8 * g8 + j];
          if (p.final_act == GTB_ACT_RELU) x = fmaxf(x, 0.f);
          else if (p.final_act == GTB_ACT_SIGMOID_AFFINE) x = p.act_eps + (1.f - 2.f * p.act_eps) * (1.f / (1.f + expf(-x)));
          v[j] = x;
        }
        *reinterpret_cast<float4*>(osl + slot_off(r, 2 * g8)) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(osl + slot_off(r, 2 * g8 + 1)) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    tc_fence_before_sync();
    __syncthreads();

    // ---------------- residual / scale, coalesced (scattered) row stores
    const int N = p.ntrue[last];
    const bool want_aggr = p.aggr != nullptr;
    const float oscale = p.out_scale ? __ldg(p.out_scale) : 1.f;
    const bool touch = p.res != nullptr || p.res_b != 1.f || p.out_scale != nullptr;
    const bool vec = (N & 3) == 0 && (p.out == nullptr || ((p.out_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)) &&
                     (p.res == nullptr || ((p.res_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.res) & 15) == 0));
    if (p.out != nullptr || touch) {
      if (vec) {
        const int c4n = N >> 2;
        for (int i = tid; i < rows_here * c4n; i += TC_NT) {
          const int rr = i / c4n, c = i - rr * c4n;
          float4* sp = reinterpret_cast<float4*>(osl + slot_off(rr, c));
          float4 v = *sp;
          if (touch) {
            v.x *= p.res_b; v.y *= p.res_b; v.z *= p.res_b; v.w *= p.res_b;
            if (p.res) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(p.res + (size_t)(row0 + rr) * p.res_ld) + c);
              v.x = fmaf(p.res_a, q.x, v.x); v.y = fmaf(p.res_a, q.y, v.y);
              v.z = fmaf(p.res_a, q.z, v.z); v.w = fmaf(p.res_a, q.w, v.w);
            }
            v.x *= oscale; v.y *= oscale; v.z *= oscale; v.w *= oscale;
            if (want_aggr) *sp = v;
          }
          if (p.out) *(reinterpret_cast<float4*>(p.out + (size_t)orow[rr] * p.out_ld) + c) = v;
        }
      } else {
        for (int i = tid; i < rows_here * N; i += TC_NT) {
          const int rr = i / N, n = i - rr * N;
          float* sp = reinterpret_cast<float*>(osl + slot_off(rr, n >> 2)) + (n & 3);
          float v = *sp;
          if (touch) {
            v *= p.res_b;
            if (p.res) v = fmaf(p.res_a, __ldg(p.res + (size_t)(row0 + rr) * p.res_ld + n), v);
            v *= oscale;
            if (want_aggr) *sp = v;
          }
          if (p.out) p.out[(size_t)orow[rr] * p.out_ld + n] = v;
        }
      }
    }

    // ---------------- in-tile segmented sum by destination (rows are destination-sorted): a run
    // covering a node's whole CSR range is stored, partial runs (tile / part boundaries) are
    // added atomically
    if (want_aggr) {
      if (touch) __syncthreads();
      constexpr int RP = 16;
      const int n_parts = (rows_here + RP - 1) / RP;
      for (int it = tid; it < N * n_parts; it += TC_NT) {
        const int part = it / N, c = it - part * N;
        const int r_beg = part * RP, r_end = min(r_beg + RP, rows_here);
        int cur = segs[r_beg];
        int g_start = r_beg;
        float sum = 0.f;
        for (int rr = r_beg; rr <= r_end; ++rr) {
          const int sg = (rr < r_end) ? segs[rr] : -2;
          if (sg != cur) {
            const int64_t gs = row0 + g_start, ge = row0 + rr;
            float* dst = p.aggr + (size_t)cur * p.aggr_ld + c;
            if (__ldg(p.rowptr + cur) == gs && __ldg(p.rowptr + cur + 1) == ge) *dst = sum;
            else atomicAdd(dst, sum);
            cur = sg;
            g_start = rr;
            sum = 0.f;
          }
          if (rr < r_end) sum += *(reinterpret_cast<const float*>(osl + slot_off(rr, c >> 2)) + (c & 3));
        }
      }
    }
    __syncthreads();
    tc_issue_item(p, issued, slots, tid);  // the output slot is free again
    ++issued;
  }

  cp_async_wait_pending(0);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
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
  for (int l = 0; l < d.n_layers; ++l) {
    p.kpad[l] = L.kpad[l];
    p.npad[l] = L.npad[l];
    p.ntrue[l] = L.ntrue[l];
    p.w_off[l][0] = L.w_off[l][0];
    p.w_off[l][1] = L.w_off[l][1];
    p.b_off[l] = L.b_off[l];
  }
  p.w_bytes = L.total_bytes;
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
  p.out = d.out;
  p.out_index = d.out_index;
  p.out_ld = d.out_ld;
  p.aggr = d.aggr;
  p.aggr_ld = d.aggr_ld;
  p.seg_id = d.seg_id;
  p.rowptr = d.rowptr;
  const size_t fixed = ((size_t)L.total_bytes + 15) / 16 * 16 + TC_MISC;
  int ring = (int)((TC_SMEM_MAX - fixed) / TC_SLOT);
  if (ring > 4) ring = 4;
  GTB_REQUIRE(ring >= 1, GTB_ERR_UNSUPPORTED_DIM, "gtb_fused_mlp_f32 (tcgen05): weights do not fit in shared memory");
  p.ring = ring;
  if (d.n_rows == 0) return GTB_OK;
  const size_t smem = fixed + (size_t)ring * TC_SLOT;
  static bool configured = false;  // one process drives one GPU (one rank per device)
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fused_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(fused_mlp_tc)");
    configured = true;
  }
  const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
  fused_mlp_tc_kernel<<<grid, TC_NT, smem, st>>>(p);
  GTB_CHECK_LAUNCH("fused_mlp_tc_kernel");
  return GTB_OK;
}

int tc_timeout_flag(int* out) {
  return check_cuda(cudaMemcpyFromSymbol(out, g_tc_timeout, sizeof(int)), "cudaMemcpyFromSymbol(g_tc_timeout)");
}

}  // namespace gtb
