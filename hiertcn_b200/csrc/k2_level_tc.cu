// K2 training levels on the tensor cores: one level of the conv stack per launch, the same operation as k2_level_f32
// (k2_tcn_f32.cu; customized_tcn_cell.py:46-49,109-127) with fp32 rows in HBM on both sides and the products on tcgen05.
//
//   forward  (kSplit):  out[r,:] = epi( sum_tap in[r - (K-1-tap)*d, :] @ W[tap] + bias [+ sbias[slot, user, :]] )
//   backward (plain):   out[r,:] = resid[r,:] + sum_tap in[r + (K-1-tap)*d, :] @ W[tap]^T          (transposed convolution)
//
// Why not the fused bf16 stack (k2_tcn_bf16.cu) for training: bf16 activations flip the ReLU gate of ~0.2 % of the
// near-zero pre-activations, a ~4 % error in every gradient below the stack.  kSplit keeps the forward at fp32 grade on
// the tensor cores: x = x_hi + x_lo, w = w_hi + w_lo (two bf16 each, 16 mantissa bits) and
//   x w  ~=  x_hi w_hi + x_lo w_hi + x_hi w_lo          (three MMAs; the dropped x_lo w_lo term is 2^-18 relative)
// accumulated in the fp32 TMEM accumulator: ~1e-5 relative, so the saved activations and gates are those of the fp32
// stack.  The backward data gradient has no gates of its own (they come from the saved forward) and runs one plain bf16 MMA.
//
// Tile = 128 rows holding floor(128 / (L + P)) whole sequences, P = (K-1)*d zero rows in front of each (behind each for the
// transposed convolution): the causal pad of customized_tcn_cell.py:46-48 as physical rows, so a tap is the SAME shared-
// memory buffer read through a descriptor whose start address is moved by the tap's shift (no-swizzle K-major layout, rows
// uniformly 16 B apart -- as in k2_tcn_bf16.cu).  Weight tiles (bf16 [tap][hi|lo][n][k], L2-resident) stream through a
// 2-stage TMA ring.  8 worker warps stage the rows (fp32 -> bf16 hi/lo) and run the epilogue ("TMEM lane = row"), warp 8
// is the TMA producer, warp 9 issues the MMAs.
#include "train.cuh"
#include "sm100.cuh"

namespace htcn {
using namespace sm100;

namespace {

constexpr int kLvRows = 128;                       // rows per tile (MMA M)
constexpr int kLvSpare = 32;                       // (K-1)*d <= 32 rows of shift
constexpr int kLvBufRows = kLvRows + kLvSpare;
constexpr int kLvActBytes = 16 * kLvBufRows * 16;  // 16 channel chunks x rows x 16 B = 40 KB per plane
constexpr int kLvWStage = 2 * 128 * 128;           // one weight tile: [128 n][128 k] bf16 as two 64-column swizzled chunks
constexpr int kLvStages = 2;
constexpr int kLvWorkWarps = 8;                    // 4 TMEM lane quarters x 2 channel halves
constexpr int kLvThreads = 32 * (kLvWorkWarps + 2);
constexpr int kLvProducerWarp = kLvWorkWarps, kLvMmaWarp = kLvWorkWarps + 1;

struct LvSlot {
  int off, L;          // first column in [B,T], length
  int unit0;           // first tile of this slot
  int spt;             // sequences per tile
};
struct LvGeom {
  int n_slots, n_units, B, T, P;
  LvSlot slot[HTCN_MAX_SLOTS];
};

// kSplit runs one CTA per SM with TWO tile buffers (and two TMEM accumulators): the worker warps stage tile i+1 and drain
// tile i-1 while the tensor pipe works on tile i (120 MMAs, ~4 us: as long as the tile's HBM traffic).  The plain variant
// (40 MMAs per tile, worker-bound) keeps one buffer and two CTAs per SM instead.
template <bool kSplit>
struct alignas(1024) LvSmem {
  static constexpr int kBufs = kSplit ? 2 : 1;
  uint8_t w[kLvStages][kLvWStage];                 // 64 KB
  uint8_t act[kBufs][kSplit ? 2 : 1][kLvActBytes]; // [tile buffer][hi (, lo) plane]
  float bias[kDim];
  uint64_t w_full[kLvStages], w_empty[kLvStages], acc_ready[2], act_ready[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t lv_desc_act(uint32_t smem_addr) {   // see make_desc_act in k2_tcn_bf16.cu
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((kLvBufRows * 16) >> 4) << 16;   // leading byte offset: the two 16-byte K chunks of a K=16 step
  d |= (uint64_t)(128 >> 4) << 32;                 // stride byte offset: 8-row groups
  d |= (uint64_t)1 << 46;
  return d;
}

// tile row r of tile `unit`: flat source row (or -1: pad / unused), its (slot, user) index and slot
__device__ __forceinline__ void lv_row(const LvGeom& g, int unit, int r, bool anti, int& src, int& sb, int& slot) {
  int s = 0;
  while (s + 1 < g.n_slots && g.slot[s + 1].unit0 <= unit) ++s;
  const LvSlot& sl = g.slot[s];
  const int stride = sl.L + g.P;
  const int seg = r / stride, w = r % stride;
  const int t = anti ? w : w - g.P;
  const int b = (unit - sl.unit0) * sl.spt + seg;
  const bool data = seg < sl.spt && b < g.B && t >= 0 && t < sl.L;
  src = data ? b * g.T + sl.off + t : -1;
  sb = s * g.B + (b < g.B ? b : 0);
  slot = s;
}

__device__ __forceinline__ void split_bf16(float x, float& hi, float& lo) {
  hi = __bfloat162float(__float2bfloat16_rn(x));
  lo = x - hi;
}

template <bool kSplit>
__global__ void __launch_bounds__(kLvThreads, kSplit ? 1 : 2)
k2_level_tc(const __grid_constant__ CUtensorMap tmap_w, LvGeom g, LevelArgs a) {
  extern __shared__ uint8_t smem_raw[];
  auto& sm = *reinterpret_cast<LvSmem<kSplit>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_wt = a.K * (kSplit ? 2 : 1);                    // weight tiles per row tile
  const int my_tiles = (g.n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool anti = a.anti != 0;

  if (tid == 0) {
    prefetch_tmap(&tmap_w);
    for (int s = 0; s < kLvStages; ++s) {
      mbar_init(&sm.w_full[s], 1);
      mbar_init(&sm.w_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sm.acc_ready[b], 1);
      mbar_init(&sm.act_ready[b], 32 * kLvWorkWarps);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < kDim; i += kLvThreads) sm.bias[i] = a.bias ? a.bias[i] : 0.f;
  for (int i = tid; i < (int)sizeof(sm.act) / 16; i += kLvThreads) reinterpret_cast<uint4*>(sm.act)[i] = make_uint4(0, 0, 0, 0);
  constexpr int kBufs = LvSmem<kSplit>::kBufs;
  if (warp == kLvMmaWarp) tmem_alloc<128 * kBufs>(&sm.tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, sm.tmem_base, 0);

  if (warp == kLvProducerWarp) {
    // ===================== weight producer: the level's n_wt tiles, once per row tile =====================
    if (lane == 0) {
      long long n = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int j = 0; j < n_wt; ++j, ++n) {
          const int s = (int)(n % kLvStages);
          mbar_wait_relaxed(&sm.w_empty[s], (uint32_t)(((n / kLvStages) & 1) ^ 1));
          mbar_arrive_expect_tx(&sm.w_full[s], kLvWStage);
          tma_load_2d(sm.w[s], &tmap_w, 0, j * 128, &sm.w_full[s]);
          tma_load_2d(sm.w[s] + kLvWStage / 2, &tmap_w, 64, j * 128, &sm.w_full[s]);
        }
      }
    }
  } else if (warp == kLvMmaWarp) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(kLvRows, 128);
    long long n = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int tb = it % kBufs;
      const uint32_t act_hi = smem_u32(sm.act[tb][0]);
      const uint32_t act_lo = smem_u32(sm.act[tb][kSplit ? 1 : 0]);
      const uint32_t tacc = tmem + tb * 128;
      mbar_wait(&sm.act_ready[tb], (uint32_t)((it / kBufs) & 1));
      tc_fence_after_sync();
      for (int tap = 0; tap < a.K; ++tap) {
        const int shift = (a.K - 1 - tap) * a.dil;
        const uint32_t row0 = (uint32_t)(anti ? shift : kLvSpare - shift) * 16;
#pragma unroll
        for (int part = 0; part < (kSplit ? 2 : 1); ++part, ++n) {
          const int s = (int)(n % kLvStages);
          mbar_wait(&sm.w_full[s], (uint32_t)((n / kLvStages) & 1));
          tc_fence_after_sync();
          const uint32_t w_base = smem_u32(sm.w[s]);
          // descriptor low words (lv_desc_act / make_desc_k_sw128), advanced by 16-byte units per K step: rebuilding the
          // 64-bit descriptors per MMA cost the issuing warp more than the MMAs take (see k2_tcn_quad.cu)
          constexpr uint32_t kAHi = (128u >> 4) | (1u << 14), kBHi = (1024u >> 4) | (1u << 14) | (2u << 29);
          const uint32_t ah_lo = (((act_hi + row0) & 0x3FFFF) >> 4) | ((uint32_t)kLvBufRows << 16);
          const uint32_t al_lo = (((act_lo + row0) & 0x3FFFF) >> 4) | ((uint32_t)kLvBufRows << 16);
          const uint32_t b_lo = ((w_base & 0x3FFFF) >> 4) | (1u << 16);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t koff = (uint32_t)(2 * k) * kLvBufRows;
            const uint32_t bk = b_lo + (uint32_t)(k >> 2) * ((kLvWStage / 2) >> 4) + (k & 3) * 2;
            if (leader) umma_bf16_lohi(tacc, ah_lo + koff, kAHi, bk, kBHi, idesc, (tap | k | part) != 0);   // x_hi w_hi, or x_hi w_lo
            if (kSplit && part == 0) {
              if (leader) umma_bf16_lohi(tacc, al_lo + koff, kAHi, bk, kBHi, idesc, true);                  // x_lo w_hi
            }
          }
          if (leader) umma_commit(&sm.w_empty[s]);
        }
      }
      if (leader) umma_commit(&sm.acc_ready[tb]);
    }
  } else {
    // ===================== stage rows + epilogue: thread = 64 channels of one tile row =====================
    const int r = tid & 127;                                   // TMEM lane r
    const int ch = tid >> 7;                                   // channel half
    const int buf_row = (anti ? 0 : kLvSpare) + r;
    const int epi = a.conv_epilogue;
    // stage the input rows of the tile whose row of this thread is `src_` into tile buffer tb
    auto stage = [&](int tb, int src_) {
      uint8_t* my_hi = sm.act[tb][0] + buf_row * 16;           // + c * kLvBufRows * 16 for channel chunk c
      uint8_t* my_lo = sm.act[tb][kSplit ? 1 : 0] + buf_row * 16;
      const float4* p = src_ >= 0 ? reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.in) + (long long)src_ * kDim + ch * 64)
                                  : nullptr;
      float4 x[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = p ? __ldg(p + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float v[8] = {x[2 * c].x, x[2 * c].y, x[2 * c].z, x[2 * c].w, x[2 * c + 1].x, x[2 * c + 1].y, x[2 * c + 1].z, x[2 * c + 1].w};
        const int off = (ch * 8 + c) * (kLvBufRows * 16);
        if (kSplit) {
          float h[8], l[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) split_bf16(v[e], h[e], l[e]);
          *reinterpret_cast<uint4*>(my_hi + off) = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
          *reinterpret_cast<uint4*>(my_lo + off) = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
        } else {
          *reinterpret_cast<uint4*>(my_hi + off) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&sm.act_ready[tb]);
    };
    int unit = blockIdx.x;
    int src, sb, slot;
    lv_row(g, unit, r, anti, src, sb, slot);
    int nsrc = -1, nsb = 0, nslot = 0;                         // the next tile's row of this thread
    stage(0, src);
    for (int it = 0; it < my_tiles; ++it) {
      const int tb = it % kBufs;
      const uint32_t tacc = tmem + tb * 128;
      if (it + 1 < my_tiles) lv_row(g, unit + (int)gridDim.x, r, anti, nsrc, nsb, nslot);
      // two buffers: tile it+1 is staged BEFORE tile it is drained (its buffer and accumulator were released by the drain of
      // tile it-1), so the tensor pipe never waits for the workers
      if (kBufs == 2 && it + 1 < my_tiles) stage((it + 1) % kBufs, nsrc);
      // ---- epilogue
      const float* res_row = nullptr;                           // row added to the result (fp32)
      if (src >= 0) {
        if (epi == 1) res_row = reinterpret_cast<const float*>(a.in) + (long long)src * kDim;
        else if (epi == 2 || epi == 3) res_row = a.resid + (long long)src * kDim;
        else if (a.sbias) res_row = a.sbias + (long long)sb * kDim;
      }
      const float* drop_row = (a.drop && (epi == 1 || epi == 3)) ? a.drop + (long long)slot * a.drop_stride : nullptr;
      float* out_row = src >= 0 ? reinterpret_cast<float*>(a.out) + (long long)src * kDim : nullptr;
      float* aux_row = (src >= 0 && a.aux && (epi == 1 || epi == 3)) ? a.aux + (long long)src * kDim : nullptr;
      mbar_wait(&sm.acc_ready[tb], (uint32_t)((it / kBufs) & 1));
      tc_fence_after_sync();
#pragma unroll 1
      for (int cc = ch * 2; cc < ch * 2 + 2; ++cc) {            // 2 x 32 channels
        float4 rs[8];
        if (res_row) {
#pragma unroll
          for (int i = 0; i < 8; ++i) rs[i] = __ldg(reinterpret_cast<const float4*>(res_row + cc * 32) + i);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        uint32_t v[32];
        tmem_ld_32x32(tacc + ((uint32_t)((warp & 3) * 32) << 16) + cc * 32, v);
        tmem_ld_wait(v);
        if (out_row) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {                          // 8 x 4 channels
            const int c0 = cc * 32 + q * 4;
            const float rv[4] = {rs[q].x, rs[q].y, rs[q].z, rs[q].w};
            float o[4], ax[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float z = __uint_as_float(v[q * 4 + e]) + sm.bias[c0 + e];
              if (epi == 1 || epi == 3) {
                z = fmaxf(z, 0.f);                               // relu(conv + b), saved before the dropout scale
                ax[e] = z;
                if (drop_row) z *= __ldg(drop_row + c0 + e);
                z = fmaxf(z + rv[e], 0.f);                       // relu(a + residual)
              } else {
                z += rv[e];                                      // in-projection: + sbias; backward: + resid
              }
              o[e] = z;
            }
            if (aux_row) *reinterpret_cast<float4*>(aux_row + c0) = make_float4(ax[0], ax[1], ax[2], ax[3]);
            *reinterpret_cast<float4*>(out_row + c0) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      tc_fence_before_sync();
      unit += gridDim.x;
      src = nsrc; sb = nsb; slot = nslot;
      if (kBufs == 1 && it + 1 < my_tiles) stage(0, src);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kLvMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc<128 * kBufs>(tmem);
  }
}

// w fp32 [K][128][128] (TF layout [tap][cin][cout]) -> bf16 tiles [tap][hi (, lo)][n][k]:
//   forward  (w_nt = 0): n = cout, k = cin      backward (w_nt = 1: W[tap] applied transposed): n = cin, k = cout
__global__ void lv_prepare_weights(const float* __restrict__ w, int w_nt, int split, __nv_bfloat16* __restrict__ out) {
  const int tap = blockIdx.x;
  const float* src = w + (long long)tap * kDim * kDim;
  __nv_bfloat16* dst = out + (long long)tap * (split ? 2 : 1) * kDim * kDim;
  for (int i = threadIdx.x; i < kDim * kDim; i += blockDim.x) {
    const int n = i / kDim, k = i % kDim;
    const float x = w_nt ? src[n * kDim + k] : src[k * kDim + n];
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    dst[i] = hi;
    if (split) dst[kDim * kDim + i] = __float2bfloat16_rn(x - __bfloat162float(hi));
  }
}

}  // namespace

bool k2_level_tc_supported(const LevelArgs& a, const SlotTable& slots) {
  if (!a.tc_ws || a.in_bf16 || a.out_bf16 || a.in_planes > 1 || a.out_row) return false;
  if (a.K < 1 || a.K > 8 || (a.K - 1) * a.dil > kLvSpare) return false;
  if (a.conv_epilogue < 0 || a.conv_epilogue > 3) return false;
  const int P = (a.K - 1) * a.dil;
  for (int s = 0; s < slots.n; ++s) {
    const int L = slots.off[s + 1] - slots.off[s];
    if (L > 0 && L + P > kLvRows) return false;              // long sequences: the streaming fused kernel / FFMA levels
  }
  return true;
}

int32_t k2_level_tc_launch(const LevelArgs& a, const SlotTable& slots, cudaStream_t st) {
  const bool split = a.tc_split != 0;
  LvGeom g{};
  g.n_slots = slots.n; g.B = a.B; g.T = a.T; g.P = (a.K - 1) * a.dil;
  int units = 0;
  for (int s = 0; s < slots.n; ++s) {
    LvSlot& sl = g.slot[s];
    sl.off = slots.off[s];
    sl.L = slots.off[s + 1] - slots.off[s];
    sl.unit0 = units;
    sl.spt = sl.L > 0 ? kLvRows / (sl.L + g.P) : 0;
    if (sl.spt > 0) units += (a.B + sl.spt - 1) / sl.spt;
  }
  g.n_units = units;
  if (units == 0) return HTCN_OK;
  __nv_bfloat16* w_bf16 = reinterpret_cast<__nv_bfloat16*>(a.tc_ws);
  const int n_wt = a.K * (split ? 2 : 1);
  lv_prepare_weights<<<a.K, 256, 0, st>>>(a.w, a.w_nt, split ? 1 : 0, w_bf16);
  HTCN_LAUNCH_CHECK("lv_prepare_weights");
  CUtensorMap tw;
  int32_t rc = make_tmap_bf16(&tw, w_bf16, (uint64_t)n_wt * kDim, kDim, kDim, 64, 128, 128);
  if (rc) return rc;
  if (split) {
    const size_t smem = sizeof(LvSmem<true>) + 1024;
    const int grid = units < 148 ? units : 148;
    HTCN_CUDA(cudaFuncSetAttribute(k2_level_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k2_level_tc<true><<<grid, kLvThreads, smem, st>>>(tw, g, a);
  } else {
    const size_t smem = sizeof(LvSmem<false>) + 1024;
    const int grid = units < 2 * 148 ? units : 2 * 148;
    HTCN_CUDA(cudaFuncSetAttribute(k2_level_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k2_level_tc<false><<<grid, kLvThreads, smem, st>>>(tw, g, a);
  }
  HTCN_LAUNCH_CHECK("k2_level_tc");
  return HTCN_OK;
}

}  // namespace htcn
