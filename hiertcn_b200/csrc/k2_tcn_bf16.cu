// K2 (bf16 tier): in-projection + dilated causal conv residual stack as a tcgen05 implicit GEMM, all levels
// fused on chip.  Replaces model_tcn.py:35 + customized_tcn_cell.py:46-49,109-127,147-161 (see k2_tcn_f32.cu for
// the op-level description; the arithmetic here is bf16 operands, fp32 accumulate, activations rounded to bf16
// between levels -- the roundings oracle/hiertcn_oracle.py's "bf16" mode mirrors).
//
// Work unit = one 128-row tile (MMA M = 128).  The rows of a tile are positions of one or more sequences laid
// out WITH their causal zero padding as physical rows:
//   short sequences (L + P <= 128, P = (K-1)*d_max): floor(128/(L+P)) sequences per tile, each preceded by P
//     zero rows -- the left pad of customized_tcn_cell.py:46-48 -- which also isolates neighbours;
//   long sequences (config 3, L = 256): a CTA STREAMS a sequence in chunks of 128 new positions.  What chunk c+1 needs
//     of chunk c is, per level, the last (K-1)*d rows of that level's INPUT; the owner threads of the tile's last 32
//     rows park them in a per-CTA global scratch (L2-resident, double-buffered by chunk parity) and cp.async them back
//     into the 32 spare rows in front of the tile while the epilogue math runs.  No receptive-field halo is recomputed
//     (the first version recomputed RF-1 = 60 of every 128 rows at config 3: 47 % of its MMAs).
// The activation tile lives in shared memory in the NO-SWIZZLE K-major UMMA layout with 8-row core matrices made
// contiguous (SBO = 128 B): row r / 16-byte channel chunk c sits at c*ROWS*16 + r*16, i.e. rows are uniformly
// 16 B apart, so conv tap k of a level with dilation d is the SAME buffer addressed through a descriptor whose
// start address is moved back by (K-1-k)*d rows.  No im2col, no per-tap copies: 5 taps = 5 descriptors.
// Weights (bf16 [tap][cout][cin], 32 KB per tap, L2-resident) stream through a 2-stage TMA ring (128B swizzle).
// Per layer: K x 8 tcgen05.mma (M=128, N=128, K=16) into one 128-column TMEM accumulator, then the 4 epilogue
// warps (TMEM lane = row) apply bias / relu / residual / relu, round to bf16 and write the next layer's operand
// in place (zero rows stay zero).  Two CTAs fit per SM (110 KB smem, 128 TMEM columns each) so one CTA's
// epilogue overlaps the other's MMAs.
#include <cstdlib>

#include "k2_tcn.cuh"

namespace htcn {
using namespace sm100;

// the 8 MMAs (K = 128) of one weight tile on one operand tile: the descriptors' low words advance by 16-byte units (the
// 64-bit rebuild per MMA -- shift, mask, two ORs per operand -- kept the issuing warp, not the tensor pipe, busy; see
// k2_tcn_quad.cu)
__device__ __forceinline__ void k2_issue_tile(bool leader, uint32_t tacc, uint32_t a_base, uint32_t w_base, uint32_t idesc,
                                              bool accumulate_first) {
  constexpr uint32_t kAHi = (128u >> 4) | (1u << 14), kBHi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t a_lo = ((a_base & 0x3FFFF) >> 4) | ((uint32_t)kRows << 16);
  const uint32_t b_lo = ((w_base & 0x3FFFF) >> 4) | (1u << 16);
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (leader)
      umma_bf16_lohi(tacc, a_lo + (uint32_t)(2 * k) * kRows, kAHi, b_lo + (uint32_t)(k >> 2) * ((kWStageBytes / 2) >> 4) + (k & 3) * 2,
                     kBHi, idesc, accumulate_first || k != 0);
}

struct alignas(1024) K2Smem {
  uint8_t w[kWStages][kWStageBytes];     // 64 KB
  uint8_t act[kActBytes];                // 40 KB
  float bias[HTCN_MAX_LEVELS][kDim];
  uint64_t w_full[kWStages], w_empty[kWStages], acc_ready, act_ready;
  uint32_t tmem_base;
};


// kPair: the CTAs run as clusters of 2 that walk the SAME weight sequence in lock step; each CTA fetches one of the two
// 64-column chunks of a weight tile and TMA-multicasts it into both CTAs' rings, so every weight byte crosses L2 -> SM once
// per CTA pair instead of once per CTA (the kernel sat at the L2 read ceiling: 352 KB of weights per 128-row tile).
// A ring stage is refilled only when BOTH CTAs' MMAs have released it (tcgen05.commit multicast onto both w_empty barriers).
// kStream: some slot is longer than a tile (chunked streaming with the per-level hand-over); kAux: anything beyond the plain
// inference stack -- down-sample levels, training dropout, saved activations.  Compile-time switches: the plain short-sequence
// kernel (the hier hot path) keeps its register budget (the all-in-one version spilled 264 B per thread and ran 22 % slower).
template <bool kPair, bool kStream, bool kAux>
__global__ void __launch_bounds__(kK2Threads, 2)
k2_tcn_bf16(const __grid_constant__ CUtensorMap tmap_w, K2Geom g, const __nv_bfloat16* __restrict__ xe,
            const float* __restrict__ sbias, const float* __restrict__ bias_all /*[n_levels][128]*/,
            const float* __restrict__ ds_bias_all /*[n_levels][128], read only for the levels of g.ds_mask*/,
            const int* __restrict__ out_row, __nv_bfloat16* __restrict__ hout,
            __nv_bfloat16* __restrict__ h_save /* [(n_levels+1)][B*T][128] every layer's output, or NULL */,
            __nv_bfloat16* __restrict__ a_save /* [n_levels][B*T][128] relu(conv + b) before the residual, or NULL */,
            uint8_t* __restrict__ hist /* [gridDim.x][2][n_levels][kHistBytes] parked level inputs of streamed sequences */,
            const float* __restrict__ drop /* [S][n_levels][128] training dropout scales (customized_tcn_cell.py:100,119) or NULL */) {
  extern __shared__ uint8_t smem_raw[];
  auto& sm = *reinterpret_cast<K2Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_layers = g.n_levels + 1;                       // layer 0 = in-projection
  // kPair: both CTAs of a pair walk the same weight sequence in lock step, so both run the larger of their two tile
  // counts; the surplus tiles of the other one are dummy tiles (all rows zero)
  int my_tiles = cta_tile_count(g, (int)blockIdx.x, (int)gridDim.x);
  if (kPair) my_tiles = max(my_tiles, cta_tile_count(g, (int)blockIdx.x ^ 1, (int)gridDim.x));
  const uint32_t crank = kPair ? cluster_ctarank() : 0u;

  if (tid == 0) {
    prefetch_tmap(&tmap_w);
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(&sm.w_full[s], 1);
      mbar_init(&sm.w_empty[s], kPair ? 2 : 1);
    }
    mbar_init(&sm.acc_ready, 1);
    mbar_init(&sm.act_ready, 32 * kK2EpiWarps);
    fence_barrier_init();
  }
  for (int i = tid; i < g.n_levels * kDim; i += kK2Threads) sm.bias[i / kDim][i % kDim] = bias_all[i];
  for (int i = tid; i < kActBytes / 16; i += kK2Threads) reinterpret_cast<uint4*>(sm.act)[i] = make_uint4(0, 0, 0, 0);
  if (warp == kK2MmaWarp) {
    if (kAux && g.ds_mask) tmem_alloc<256>(&sm.tmem_base);    // second accumulator: the down-sample residual
    else tmem_alloc<128>(&sm.tmem_base);
  }
  tc_fence_before_sync();
  if (kPair) cluster_sync_all();                  // the peer's barriers are initialised before any multicast lands on them
  else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, sm.tmem_base, 0);

  if (warp == kK2ProducerWarp) {
    // ===================== weight producer: [W_in, L0 taps, L1 taps, ...] per tile, 2-stage ring =====================
    if (lane == 0) {
      const int per_tile = 1 + g.n_levels * g.K + (kAux ? __popc(g.ds_mask) : 0);
      long long n = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int j = 0; j < per_tile; ++j, ++n) {
          const int s = (int)(n % kWStages);
          mbar_wait_relaxed(&sm.w_empty[s], (uint32_t)(((n / kWStages) & 1) ^ 1));
          mbar_arrive_expect_tx(&sm.w_full[s], kWStageBytes);
          if (kPair) {                                           // this CTA's chunk, into both rings
            tma_load_2d_multicast(sm.w[s] + crank * (kWStageBytes / 2), &tmap_w, 64 * (int)crank, j * 128, &sm.w_full[s], 0b11);
          } else {
            tma_load_2d(sm.w[s], &tmap_w, 0, j * 128, &sm.w_full[s]);
            tma_load_2d(sm.w[s] + kWStageBytes / 2, &tmap_w, 64, j * 128, &sm.w_full[s]);
          }
        }
      }
    }
  } else if (warp == kK2MmaWarp) {
    // ===================== MMA issuer =====================
    // the whole warp runs the (warp-uniform) waits and descriptor arithmetic, one elected lane issues: under an
    // `if (lane == 0)` branch every tcgen05.mma is wrapped in an ELECT / R2UR.BROADCAST waterfall (~100 clk per MMA)
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(kTR, 128);
      const uint32_t act0 = smem_u32(sm.act);
      long long n = 0, n_act = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int layer = 0; layer < n_layers; ++layer, ++n_act) {
          mbar_wait(&sm.act_ready, (uint32_t)(n_act & 1));       // operand tile written + fenced by the epilogue warps
          tc_fence_after_sync();
          const int taps = layer == 0 ? 1 : g.K;
          const int dil = layer == 0 ? 1 : (1 << (layer - 1));
          for (int tap = 0; tap < taps; ++tap, ++n) {
            const int s = (int)(n % kWStages);
            mbar_wait(&sm.w_full[s], (uint32_t)((n / kWStages) & 1));
            tc_fence_after_sync();
            const int shift = (taps - 1 - tap) * dil;            // rows back in time (customized_tcn_cell.py:46-49)
            const uint32_t a_base = act0 + (uint32_t)(kMaxSpare - shift) * 16;
            const uint32_t w_base = smem_u32(sm.w[s]);
            k2_issue_tile(leader, tmem, a_base, w_base, idesc, tap != 0);
            if (leader) {
              if (kPair) umma_commit_mc(&sm.w_empty[s], 0b11);
              else umma_commit(&sm.w_empty[s]);
            }
          }
          if (kAux && layer > 0 && ((g.ds_mask >> (layer - 1)) & 1u)) {  // res = in @ W_ds: unshifted rows, second accumulator
            const int s = (int)(n % kWStages);
            mbar_wait(&sm.w_full[s], (uint32_t)((n / kWStages) & 1));
            tc_fence_after_sync();
            const uint32_t a_base = act0 + (uint32_t)kMaxSpare * 16;
            const uint32_t w_base = smem_u32(sm.w[s]);
            k2_issue_tile(leader, tmem + 128, a_base, w_base, idesc, false);
            if (leader) {
              if (kPair) umma_commit_mc(&sm.w_empty[s], 0b11);
              else umma_commit(&sm.w_empty[s]);
            }
            ++n;
          }
          if (leader) umma_commit(&sm.acc_ready);
        }
      }
    }
  } else {
    // ===================== loader + epilogue: thread = 64 channels (half ch) of one tile row =====================
    // (with 4 epilogue warps -- one per scheduler -- the dependent FADD/FMNMX chains of the epilogue issued at ~4.5 clk per
    // instruction and the MMA warp spent 2/3 of its time waiting for act_ready: ncu source view, profiles/r1_k2_ncu.txt)
    const int r = tid & 127;                                    // TMEM lane r (warp w may touch lanes 32(w%4)..+31)
    const int ch = tid >> 7;                                    // channel half: 64*ch .. 64*ch+63
    const uint32_t my_act = smem_u32(sm.act) + (kMaxSpare + r) * 16;      // + c * kRows * 16 for channel chunk c (shared window)
    long long n_acc = 0;
    int unit = blockIdx.x, chunk = 0, unit_chunks = unit < g.n_units ? g.slot[unit_slot(g, unit)].chunks : 1;
    // streaming of long sequences: the threads of the tile's last kMaxSpare rows own the hand-over to the next chunk
    const bool hist_owner = kStream && r >= kTR - kMaxSpare;
    const int hj = r - (kTR - kMaxSpare);                       // spare row / parked row of this thread
    uint8_t* my_hist = hist + (size_t)blockIdx.x * 2 * g.n_levels * kHistBytes;
    bool spare_dirty = false;                                   // the spare rows hold parked data (not the zero pad)
    for (int it = 0; it < my_tiles; ++it) {
      int src, dst, sb, slot_idx;
      bool own;
      tile_geometry(g, unit, chunk, r, out_row, src, dst, sb, own, slot_idx);
      const bool streaming = kStream && unit_chunks > 1;
      const long long RT = (long long)g.B * g.T;
      // ---- stage the input rows (bf16 Xe) into the operand layout; zero rows stay zero
      {
        const uint4* p = src >= 0 ? reinterpret_cast<const uint4*>(xe + (long long)src * kDim) : nullptr;
#pragma unroll
        for (int c = ch * 8; c < ch * 8 + 8; ++c)
          sts_u4(my_act + c * (kRows * 16), p ? __ldg(p + c) : make_uint4(0, 0, 0, 0));
      }
      fence_proxy_async_smem();
      mbar_arrive(&sm.act_ready);
      for (int layer = 0; layer < n_layers; ++layer, ++n_acc) {
        mbar_wait(&sm.acc_ready, (uint32_t)(n_acc & 1));
        tc_fence_after_sync();
        const bool last = layer == n_layers - 1;
        // The MMAs that read the spare rows have retired: refill them for the NEXT layer (conv level `layer`), whose
        // taps reach back up to kMaxSpare rows -- with the rows chunk-1 parked for that level, or with the causal zero pad.
        uint8_t* park = my_hist + ((size_t)(chunk & 1) * g.n_levels + layer) * kHistBytes;
        if (kStream && !last && hist_owner) {
          const uint8_t* prev = my_hist + ((size_t)((chunk & 1) ^ 1) * g.n_levels + layer) * kHistBytes;
          if (streaming && chunk > 0) {
#pragma unroll
            for (int c = ch * 8; c < ch * 8 + 8; ++c)
              cp_async_16(sm.act + c * (kRows * 16) + hj * 16, prev + (c * kMaxSpare + hj) * 16);
            spare_dirty = true;
          } else if (spare_dirty) {                              // back to the causal zero pad
#pragma unroll
            for (int c = ch * 8; c < ch * 8 + 8; ++c)
              sts_u4(smem_u32(sm.act) + c * (kRows * 16) + hj * 16, make_uint4(0, 0, 0, 0));
            spare_dirty = false;
          }
        }
        const uint32_t bias_s = smem_u32(sm.bias[layer > 0 ? layer - 1 : 0]);      // shared-window address of the level's bias
        const bool ds = kAux && layer > 0 && ((g.ds_mask >> (layer - 1)) & 1u);
        const float* ds_bias_l = ds ? ds_bias_all + (layer - 1) * kDim : nullptr;
        const float* drop_l = (kAux && drop && layer > 0) ? drop + ((long long)slot_idx * g.n_levels + (layer - 1)) * kDim : nullptr;
        const float* sb_row = (layer == 0 && sbias && src >= 0) ? sbias + (long long)sb * kDim : nullptr;
#pragma unroll 1
        for (int cc = ch * 2; cc < ch * 2 + 2; ++cc) {           // 2 x 32 channels
          // the per-(slot, user) bias of the in-projection: all 8 loads of this 32-channel group in flight BEFORE the TMEM
          // read (inside the q loop each 16-byte group exposed a full L2 latency: 22 % of the epilogue warps' samples)
          float4 sbv[8];
          if (sb_row) {
#pragma unroll
            for (int i = 0; i < 8; ++i) sbv[i] = __ldg(reinterpret_cast<const float4*>(sb_row + cc * 32) + i);
          }
          uint32_t v[32];
          tmem_ld_32x32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + cc * 32, v);
          tmem_ld_wait(v);
#pragma unroll
          for (int q = 0; q < 4; ++q) {                          // 4 x 8 channels = one 16-byte chunk each
            const int c = cc * 4 + q;
            const uint32_t slot = my_act + c * (kRows * 16);
            float o[8], av[8];
            if (layer == 0) {
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[q * 8 + e]);
              if (sb_row) {
                const float4 s0 = sbv[2 * q], s1 = sbv[2 * q + 1];
                o[0] += s0.x; o[1] += s0.y; o[2] += s0.z; o[3] += s0.w;
                o[4] += s1.x; o[5] += s1.y; o[6] += s1.z; o[7] += s1.w;
              }
            } else {
              float rs[8];
              if (ds) {                                          // residual = in @ W_ds + b_ds from the second accumulator
                uint32_t v2[8];
                tmem_ld_32x8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 128 + c * 8, v2);
                tmem_ld_wait(v2);
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(ds_bias_l + c * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(ds_bias_l + c * 8) + 1);
                rs[0] = __uint_as_float(v2[0]) + b0.x; rs[1] = __uint_as_float(v2[1]) + b0.y;
                rs[2] = __uint_as_float(v2[2]) + b0.z; rs[3] = __uint_as_float(v2[3]) + b0.w;
                rs[4] = __uint_as_float(v2[4]) + b1.x; rs[5] = __uint_as_float(v2[5]) + b1.y;
                rs[6] = __uint_as_float(v2[6]) + b1.z; rs[7] = __uint_as_float(v2[7]) + b1.w;
              } else {
                const uint4 res = lds_u4(slot);                  // this row's input to the level (bf16 x 8)
                rs[0] = bf16_lo(res.x); rs[1] = bf16_hi(res.x); rs[2] = bf16_lo(res.y); rs[3] = bf16_hi(res.y);
                rs[4] = bf16_lo(res.z); rs[5] = bf16_hi(res.z); rs[6] = bf16_lo(res.w); rs[7] = bf16_hi(res.w);
              }
              const float4 bv0 = lds_f4(bias_s + c * 32), bv1 = lds_f4(bias_s + c * 32 + 16);
              const float bl[8] = {bv0.x, bv0.y, bv0.z, bv0.w, bv1.x, bv1.y, bv1.z, bv1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float a = fmaxf(__uint_as_float(v[q * 8 + e]) + bl[e], 0.f);                     // relu(conv + b)
                av[e] = a;                                                                       // saved before dropout
                if (drop_l) a *= __ldg(drop_l + c * 8 + e);                                      // training only
                o[e] = fmaxf(a + rs[e], 0.f);                                                    // relu(a + residual)
              }
              if (kAux && a_save && own)
                reinterpret_cast<uint4*>(a_save + ((long long)(layer - 1) * RT + src) * kDim)[c] =
                    make_uint4(pack_bf16x2(av[0], av[1]), pack_bf16x2(av[2], av[3]), pack_bf16x2(av[4], av[5]),
                               pack_bf16x2(av[6], av[7]));
            }
            uint4 packed = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                                      pack_bf16x2(o[6], o[7]));
            if (src < 0) packed = make_uint4(0, 0, 0, 0);        // causal pad rows stay zero at every level
            if (kAux && h_save && own) reinterpret_cast<uint4*>(h_save + ((long long)layer * RT + src) * kDim)[c] = packed;
            if (!last) {
              sts_u4(slot, packed);
              // park this row of the next level's input for the sequence's next chunk (read back by this same thread)
              if (kStream && streaming && hist_owner && chunk + 1 < unit_chunks)
                *reinterpret_cast<uint4*>(park + (c * kMaxSpare + hj) * 16) = packed;
            } else if (dst >= 0) reinterpret_cast<uint4*>(hout + (long long)dst * kDim)[c] = packed;
          }
        }
        tc_fence_before_sync();
        if (!last) {
          if (kStream && hist_owner) cp_async_wait_all();        // the parked rows have landed in the spare rows
          fence_proxy_async_smem();
          mbar_arrive(&sm.act_ready);
        }
      }
      if (++chunk == unit_chunks) {
        unit += gridDim.x;
        chunk = 0;
        unit_chunks = unit < g.n_units ? g.slot[unit_slot(g, unit)].chunks : 1;
      }
    }
  }
  tc_fence_before_sync();
  if (kPair) cluster_sync_all();                  // nobody exits while the peer may still multicast into its ring
  else __syncthreads();
  if (warp == kK2MmaWarp) {
    tc_fence_after_sync();
    if (kAux && g.ds_mask) tmem_dealloc<256>(tmem);
    else tmem_dealloc<128>(tmem);
  }
}


// ---- two tile chains per CTA (the default) ---------------------------------------------------------------------------
// The kernel above keeps ONE tile per CTA and two CTAs per SM; a tile's layer is a strictly serial MMA phase (5 taps through
// a 2-deep weight ring: the refill latency of the ring, ~2000 clk per tap, not the 512 clk of its MMAs) followed by an
// epilogue phase, so each CTA keeps the tensor pipe busy ~20 % of the time (ncu: 39-41 % with two CTAs).  Here ONE CTA per SM
// runs two independent tile chains (two operand tiles, two TMEM accumulators, two groups of 8 worker warps) served by one
// MMA warp that alternates between them layer by layer, and the weight ring is FOUR stages deep (128 KB): while chain A's
// epilogue runs, the tensor pipe works on chain B, and a layer's taps are already in shared memory when its turn comes.
// Chain c of CTA b behaves like "virtual CTA" 2b + c of the single-chain kernel (units, streaming hand-over scratch).
constexpr int kDualWStages = 4;
constexpr int kK2DualThreads = 32 * (2 * kK2EpiWarps + 2);       // warps 0-7 chain 0, 8-15 chain 1, 16 TMA producer, 17 MMA
constexpr int kK2DualProducerWarp = 2 * kK2EpiWarps, kK2DualMmaWarp = 2 * kK2EpiWarps + 1;
struct alignas(1024) K2SmemDual {
  uint8_t w[kDualWStages][kWStageBytes];   // 128 KB
  uint8_t act[2][kActBytes];               // 80 KB
  float bias[HTCN_MAX_LEVELS][kDim];
  uint64_t w_full[kDualWStages], w_empty[kDualWStages], acc_ready[2], act_ready[2];
  uint32_t tmem_base;
};


template <bool kStream, bool kAux>
__global__ void __launch_bounds__(kK2DualThreads, 1)
k2_tcn_bf16_dual(const __grid_constant__ CUtensorMap tmap_w, K2Geom g, const __nv_bfloat16* __restrict__ xe,
                 const float* __restrict__ sbias, const float* __restrict__ bias_all, const float* __restrict__ ds_bias_all,
                 const int* __restrict__ out_row, __nv_bfloat16* __restrict__ hout, __nv_bfloat16* __restrict__ h_save,
                 __nv_bfloat16* __restrict__ a_save, uint8_t* __restrict__ hist, const float* __restrict__ drop) {
  extern __shared__ uint8_t smem_raw[];
  auto& sm = *reinterpret_cast<K2SmemDual*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_layers = g.n_levels + 1;                       // layer 0 = in-projection
  const int n_vc = 2 * (int)gridDim.x;                       // virtual CTAs = tile chains
  const int tiles0 = cta_tile_count(g, 2 * (int)blockIdx.x, n_vc), tiles1 = cta_tile_count(g, 2 * (int)blockIdx.x + 1, n_vc);
  const int steps0 = tiles0 * n_layers, steps1 = tiles1 * n_layers;      // layers each chain runs
  const int n_rounds = steps0 > steps1 ? steps0 : steps1;
  const bool two_acc = kAux && g.ds_mask != 0;               // second accumulator per chain: the down-sample residual

  if (tid == 0) {
    prefetch_tmap(&tmap_w);
    for (int s = 0; s < kDualWStages; ++s) {
      mbar_init(&sm.w_full[s], 1);
      mbar_init(&sm.w_empty[s], 1);
    }
    for (int c = 0; c < 2; ++c) {
      mbar_init(&sm.acc_ready[c], 1);
      mbar_init(&sm.act_ready[c], 32 * kK2EpiWarps);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < g.n_levels * kDim; i += kK2DualThreads) sm.bias[i / kDim][i % kDim] = bias_all[i];
  for (int i = tid; i < 2 * kActBytes / 16; i += kK2DualThreads) reinterpret_cast<uint4*>(sm.act)[i] = make_uint4(0, 0, 0, 0);
  if (warp == kK2DualMmaWarp) {
    if (two_acc) tmem_alloc<512>(&sm.tmem_base);
    else tmem_alloc<256>(&sm.tmem_base);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, sm.tmem_base, 0);
  const uint32_t acc_cols = two_acc ? 256u : 128u;           // TMEM columns per chain

  if (warp == kK2DualProducerWarp) {
    // ===================== weight producer: the tiles of (chain 0, layer i), (chain 1, layer i), ... =====================
    if (lane == 0) {
      long long n = 0;
      for (int i = 0; i < n_rounds; ++i) {
        int first, taps;
        bool ds;
        k2_layer_tiles(g, i % n_layers, first, taps, ds);
        const int cnt = taps + ((kAux && ds) ? 1 : 0);
        for (int cn = 0; cn < 2; ++cn) {
          if (i >= (cn ? steps1 : steps0)) continue;
          for (int j = first; j < first + cnt; ++j, ++n) {
            const int s = (int)(n % kDualWStages);
            mbar_wait_relaxed(&sm.w_empty[s], (uint32_t)(((n / kDualWStages) & 1) ^ 1));
            mbar_arrive_expect_tx(&sm.w_full[s], kWStageBytes);
            tma_load_2d(sm.w[s], &tmap_w, 0, j * 128, &sm.w_full[s]);
            tma_load_2d(sm.w[s] + kWStageBytes / 2, &tmap_w, 64, j * 128, &sm.w_full[s]);
          }
        }
      }
    }
  } else if (warp == kK2DualMmaWarp) {
    // ===================== MMA issuer: alternates between the chains layer by layer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(kTR, 128);
    long long n = 0;
    for (int i = 0; i < n_rounds; ++i) {
      const int layer = i % n_layers;
      int first, taps;
      bool ds;
      k2_layer_tiles(g, layer, first, taps, ds);
      const int dil = layer == 0 ? 1 : (1 << (layer - 1));
      for (int cn = 0; cn < 2; ++cn) {
        if (i >= (cn ? steps1 : steps0)) continue;
        const uint32_t act0 = smem_u32(sm.act[cn]);
        const uint32_t tacc = tmem + (uint32_t)cn * acc_cols;
        mbar_wait(&sm.act_ready[cn], (uint32_t)(i & 1));         // operand tile written + fenced by the chain's worker warps
        tc_fence_after_sync();
        for (int tap = 0; tap < taps; ++tap, ++n) {
          const int s = (int)(n % kDualWStages);
          mbar_wait(&sm.w_full[s], (uint32_t)((n / kDualWStages) & 1));
          tc_fence_after_sync();
          const int shift = (taps - 1 - tap) * dil;              // rows back in time (customized_tcn_cell.py:46-49)
          const uint32_t a_base = act0 + (uint32_t)(kMaxSpare - shift) * 16;
          const uint32_t w_base = smem_u32(sm.w[s]);
          k2_issue_tile(leader, tacc, a_base, w_base, idesc, tap != 0);
          if (leader) umma_commit(&sm.w_empty[s]);
        }
        if (kAux && ds) {                                        // res = in @ W_ds: unshifted rows, second accumulator
          const int s = (int)(n % kDualWStages);
          mbar_wait(&sm.w_full[s], (uint32_t)((n / kDualWStages) & 1));
          tc_fence_after_sync();
          const uint32_t a_base = act0 + (uint32_t)kMaxSpare * 16;
          const uint32_t w_base = smem_u32(sm.w[s]);
          k2_issue_tile(leader, tacc + 128, a_base, w_base, idesc, false);
          if (leader) umma_commit(&sm.w_empty[s]);
          ++n;
        }
        if (leader) umma_commit(&sm.acc_ready[cn]);
      }
    }
  } else {
    const int cn = warp >> 3;                                   // tile chain of this worker warp
    const int vc = 2 * (int)blockIdx.x + cn;                    // its virtual CTA
    const int tiles_g = cn ? tiles1 : tiles0;
    const uint32_t tmem_c = tmem + (uint32_t)cn * acc_cols;
    // ===================== loader + epilogue: thread = 64 channels (half ch) of one tile row =====================
    // (with 4 epilogue warps -- one per scheduler -- the dependent FADD/FMNMX chains of the epilogue issued at ~4.5 clk per
    // instruction and the MMA warp spent 2/3 of its time waiting for act_ready: ncu source view, profiles/r1_k2_ncu.txt)
    const int tg = tid - 256 * cn;                                 // thread index inside the chain's worker group
    const int r = tg & 127;                                    // TMEM lane r (warp w may touch lanes 32(w%4)..+31)
    const int ch = tg >> 7;                                     // channel half: 64*ch .. 64*ch+63
    const uint32_t my_act = smem_u32(sm.act[cn]) + (kMaxSpare + r) * 16;  // + c * kRows * 16 for channel chunk c (shared window)
    long long n_acc = 0;
    int unit = vc, chunk = 0, unit_chunks = unit < g.n_units ? g.slot[unit_slot(g, unit)].chunks : 1;
    // streaming of long sequences: the threads of the tile's last kMaxSpare rows own the hand-over to the next chunk
    const bool hist_owner = kStream && r >= kTR - kMaxSpare;
    const int hj = r - (kTR - kMaxSpare);                       // spare row / parked row of this thread
    uint8_t* my_hist = hist + (size_t)vc * 2 * g.n_levels * kHistBytes;
    bool spare_dirty = false;                                   // the spare rows hold parked data (not the zero pad)
    for (int it = 0; it < tiles_g; ++it) {
      int src, dst, sb, slot_idx;
      bool own;
      tile_geometry(g, unit, chunk, r, out_row, src, dst, sb, own, slot_idx);
      const bool streaming = kStream && unit_chunks > 1;
      const long long RT = (long long)g.B * g.T;
      // ---- stage the input rows (bf16 Xe) into the operand layout; zero rows stay zero
      {
        const uint4* p = src >= 0 ? reinterpret_cast<const uint4*>(xe + (long long)src * kDim) : nullptr;
#pragma unroll
        for (int c = ch * 8; c < ch * 8 + 8; ++c)
          sts_u4(my_act + c * (kRows * 16), p ? __ldg(p + c) : make_uint4(0, 0, 0, 0));
      }
      fence_proxy_async_smem();
      mbar_arrive(&sm.act_ready[cn]);
      for (int layer = 0; layer < n_layers; ++layer, ++n_acc) {
        mbar_wait(&sm.acc_ready[cn], (uint32_t)(n_acc & 1));
        tc_fence_after_sync();
        const bool last = layer == n_layers - 1;
        // The MMAs that read the spare rows have retired: refill them for the NEXT layer (conv level `layer`), whose
        // taps reach back up to kMaxSpare rows -- with the rows chunk-1 parked for that level, or with the causal zero pad.
        uint8_t* park = my_hist + ((size_t)(chunk & 1) * g.n_levels + layer) * kHistBytes;
        if (kStream && !last && hist_owner) {
          const uint8_t* prev = my_hist + ((size_t)((chunk & 1) ^ 1) * g.n_levels + layer) * kHistBytes;
          if (streaming && chunk > 0) {
#pragma unroll
            for (int c = ch * 8; c < ch * 8 + 8; ++c)
              cp_async_16(sm.act[cn] + c * (kRows * 16) + hj * 16, prev + (c * kMaxSpare + hj) * 16);
            spare_dirty = true;
          } else if (spare_dirty) {                              // back to the causal zero pad
#pragma unroll
            for (int c = ch * 8; c < ch * 8 + 8; ++c)
              sts_u4(smem_u32(sm.act[cn]) + c * (kRows * 16) + hj * 16, make_uint4(0, 0, 0, 0));
            spare_dirty = false;
          }
        }
        const uint32_t bias_s = smem_u32(sm.bias[layer > 0 ? layer - 1 : 0]);      // shared-window address of the level's bias
        const bool ds = kAux && layer > 0 && ((g.ds_mask >> (layer - 1)) & 1u);
        const float* ds_bias_l = ds ? ds_bias_all + (layer - 1) * kDim : nullptr;
        const float* drop_l = (kAux && drop && layer > 0) ? drop + ((long long)slot_idx * g.n_levels + (layer - 1)) * kDim : nullptr;
        const float* sb_row = (layer == 0 && sbias && src >= 0) ? sbias + (long long)sb * kDim : nullptr;
#pragma unroll 1
        for (int cc = ch * 2; cc < ch * 2 + 2; ++cc) {           // 2 x 32 channels
          // the per-(slot, user) bias of the in-projection: all 8 loads of this 32-channel group in flight BEFORE the TMEM
          // read (inside the q loop each 16-byte group exposed a full L2 latency: 22 % of the epilogue warps' samples)
          float4 sbv[8];
          if (sb_row) {
#pragma unroll
            for (int i = 0; i < 8; ++i) sbv[i] = __ldg(reinterpret_cast<const float4*>(sb_row + cc * 32) + i);
          }
          uint32_t v[32];
          tmem_ld_32x32(tmem_c + ((uint32_t)((warp & 3) * 32) << 16) + cc * 32, v);
          tmem_ld_wait(v);
#pragma unroll
          for (int q = 0; q < 4; ++q) {                          // 4 x 8 channels = one 16-byte chunk each
            const int c = cc * 4 + q;
            const uint32_t slot = my_act + c * (kRows * 16);
            float o[8], av[8];
            if (layer == 0) {
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[q * 8 + e]);
              if (sb_row) {
                const float4 s0 = sbv[2 * q], s1 = sbv[2 * q + 1];
                o[0] += s0.x; o[1] += s0.y; o[2] += s0.z; o[3] += s0.w;
                o[4] += s1.x; o[5] += s1.y; o[6] += s1.z; o[7] += s1.w;
              }
            } else {
              float rs[8];
              if (ds) {                                          // residual = in @ W_ds + b_ds from the second accumulator
                uint32_t v2[8];
                tmem_ld_32x8(tmem_c + ((uint32_t)((warp & 3) * 32) << 16) + 128 + c * 8, v2);
                tmem_ld_wait(v2);
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(ds_bias_l + c * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(ds_bias_l + c * 8) + 1);
                rs[0] = __uint_as_float(v2[0]) + b0.x; rs[1] = __uint_as_float(v2[1]) + b0.y;
                rs[2] = __uint_as_float(v2[2]) + b0.z; rs[3] = __uint_as_float(v2[3]) + b0.w;
                rs[4] = __uint_as_float(v2[4]) + b1.x; rs[5] = __uint_as_float(v2[5]) + b1.y;
                rs[6] = __uint_as_float(v2[6]) + b1.z; rs[7] = __uint_as_float(v2[7]) + b1.w;
              } else {
                const uint4 res = lds_u4(slot);                  // this row's input to the level (bf16 x 8)
                rs[0] = bf16_lo(res.x); rs[1] = bf16_hi(res.x); rs[2] = bf16_lo(res.y); rs[3] = bf16_hi(res.y);
                rs[4] = bf16_lo(res.z); rs[5] = bf16_hi(res.z); rs[6] = bf16_lo(res.w); rs[7] = bf16_hi(res.w);
              }
              const float4 bv0 = lds_f4(bias_s + c * 32), bv1 = lds_f4(bias_s + c * 32 + 16);
              const float bl[8] = {bv0.x, bv0.y, bv0.z, bv0.w, bv1.x, bv1.y, bv1.z, bv1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float a = fmaxf(__uint_as_float(v[q * 8 + e]) + bl[e], 0.f);                     // relu(conv + b)
                av[e] = a;                                                                       // saved before dropout
                if (drop_l) a *= __ldg(drop_l + c * 8 + e);                                      // training only
                o[e] = fmaxf(a + rs[e], 0.f);                                                    // relu(a + residual)
              }
              if (kAux && a_save && own)
                reinterpret_cast<uint4*>(a_save + ((long long)(layer - 1) * RT + src) * kDim)[c] =
                    make_uint4(pack_bf16x2(av[0], av[1]), pack_bf16x2(av[2], av[3]), pack_bf16x2(av[4], av[5]),
                               pack_bf16x2(av[6], av[7]));
            }
            uint4 packed = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                                      pack_bf16x2(o[6], o[7]));
            if (src < 0) packed = make_uint4(0, 0, 0, 0);        // causal pad rows stay zero at every level
            if (kAux && h_save && own) reinterpret_cast<uint4*>(h_save + ((long long)layer * RT + src) * kDim)[c] = packed;
            if (!last) {
              sts_u4(slot, packed);
              // park this row of the next level's input for the sequence's next chunk (read back by this same thread)
              if (kStream && streaming && hist_owner && chunk + 1 < unit_chunks)
                *reinterpret_cast<uint4*>(park + (c * kMaxSpare + hj) * 16) = packed;
            } else if (dst >= 0) reinterpret_cast<uint4*>(hout + (long long)dst * kDim)[c] = packed;
          }
        }
        tc_fence_before_sync();
        if (!last) {
          if (kStream && hist_owner) cp_async_wait_all();        // the parked rows have landed in the spare rows
          fence_proxy_async_smem();
          mbar_arrive(&sm.act_ready[cn]);
        }
      }
      if (++chunk == unit_chunks) {
        unit += n_vc;
        chunk = 0;
        unit_chunks = unit < g.n_units ? g.slot[unit_slot(g, unit)].chunks : 1;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kK2DualMmaWarp) {
    tc_fence_after_sync();
    if (two_acc) tmem_dealloc<512>(tmem);
    else tmem_dealloc<256>(tmem);
  }
}

// weights f32 [tap][cin][cout] (TF layout, customized_convolution_layer.py:137) -> bf16 [tap][cout][cin]
// tile_src[j]: the [128 cin][128 cout] fp32 source of weight tile j, in the order the kernel consumes them:
// in-projection, then per level its K taps and, if it has one, the down-sample Dense kernel
constexpr int kK2MaxWeightTiles = 1 + HTCN_MAX_LEVELS * 9;
constexpr int kK2PtrTableBytes = 640;            // >= kK2MaxWeightTiles pointers, keeps the bias arrays 16-byte aligned
static_assert(kK2MaxWeightTiles * 8 <= kK2PtrTableBytes, "pointer table");
// One launch prepares everything the stack kernels read from the scratch: the bf16 weight tiles and the level biases.  The
// source pointers travel as a kernel argument (a pointer table in device memory cost a pageable host-to-device memcpy plus
// one device-to-device memcpy per bias vector on every call: ~0.1 ms of launches around a 0.43 ms kernel).
struct K2PrepArgs {
  const float* tile[kK2MaxWeightTiles];          // [128 cin][128 cout] fp32 sources, consumption order
  const float* bias[HTCN_MAX_LEVELS];            // conv biases
  const float* ds_bias[HTCN_MAX_LEVELS];         // down-sample biases (NULL: zeros / not a down-sample level)
  int n_tiles, n_levels;
  unsigned ds_mask;
};
__global__ void k2_prepare_weights(const __grid_constant__ K2PrepArgs a, __nv_bfloat16* __restrict__ out, float* __restrict__ bias_dev,
                                   float* __restrict__ ds_bias_dev) {
  const int j = blockIdx.x;
  if (j == a.n_tiles) {                            // the extra block: biases
    for (int i = threadIdx.x; i < a.n_levels * kDim; i += blockDim.x) {
      const int l = i / kDim, c = i % kDim;
      bias_dev[i] = a.bias[l][c];
      if ((a.ds_mask >> l) & 1u) ds_bias_dev[i] = a.ds_bias[l] ? a.ds_bias[l][c] : 0.f;
    }
    return;
  }
  const float* src = a.tile[j];
  for (int i = threadIdx.x; i < kDim * kDim; i += blockDim.x) {
    const int cout = i / kDim, cin = i % kDim;
    out[(long long)j * kDim * kDim + i] = __float2bfloat16_rn(src[cin * kDim + cout]);
  }
}

int32_t tcn_forward_bf16(const void* xe, int xe_dtype, const float* w_in_x, const float* sbias,
                         const float* const* conv_w, const float* const* conv_b, const float* const* ds_w,
                         const float* const* ds_b, int n_levels, int K,
                         const SlotTable& slots, int B, int T, const int* out_row, void* hout, int hout_dtype,
                         float* scratch, cudaStream_t st, void* h_save, void* a_save, const float* drop) {
  if (xe_dtype != HTCN_BF16 || hout_dtype != HTCN_BF16) {
    set_error("tcn_forward(bf16): xe and hout must be bf16");
    return HTCN_ERR_INVALID;
  }
  if (!scratch) {
    set_error("tcn_forward(bf16): scratch is required (%d bytes for the bf16 weight tiles)",
              (int)HTCN_TCN_SCRATCH_BYTES(n_levels, K));
    return HTCN_ERR_INVALID;
  }
  const int P = n_levels > 0 ? (K - 1) * (1 << (n_levels - 1)) : 0;
  if (P > kMaxSpare) {
    set_error("tcn_forward(bf16): (K-1)*2^(levels-1) = %d rows of causal shift exceed the fused kernel's %d; use HTCN_F32",
              P, kMaxSpare);
    return HTCN_ERR_UNSUPPORTED;
  }
  K2Geom g{};
  g.n_slots = slots.n; g.B = B; g.T = T; g.K = K; g.n_levels = n_levels; g.P = P;
  int units = 0;
  for (int s = 0; s < slots.n; ++s) {
    K2Slot& sl = g.slot[s];
    sl.off = slots.off[s];
    sl.L = slots.off[s + 1] - slots.off[s];
    sl.unit0 = units;
    if (sl.L + P <= kTR) {               // short: several zero-padded sequences per tile
      sl.seq_per_tile = kTR / (sl.L + P);
      sl.chunks = 1;
      units += (B + sl.seq_per_tile - 1) / sl.seq_per_tile;
    } else {                             // long: one sequence per unit, streamed in chunks of 128 positions
      sl.seq_per_tile = 0;
      sl.chunks = (sl.L + kTR - 1) / kTR;
      units += B;
    }
  }
  g.n_units = units;
  // scratch layout: [bf16 weight tiles (HTCN_TCN_SCRATCH_BYTES reserves K+1 per level)][tile source pointers][conv biases]
  // [down-sample biases]
  K2PrepArgs pa{};
  int n_wt = 0;
  pa.tile[n_wt++] = w_in_x;
  for (int l = 0; l < n_levels; ++l) {
    for (int tap = 0; tap < K; ++tap) pa.tile[n_wt++] = conv_w[l] + (long long)tap * kDim * kDim;
    pa.bias[l] = conv_b[l];
    if (ds_w && ds_w[l]) {
      g.ds_mask |= 1u << l;
      pa.tile[n_wt++] = ds_w[l];
      pa.ds_bias[l] = (ds_b && ds_b[l]) ? ds_b[l] : nullptr;
    }
  }
  pa.n_tiles = n_wt; pa.n_levels = n_levels; pa.ds_mask = g.ds_mask;
  // scratch layout: [bf16 weight tiles (HTCN_TCN_SCRATCH_BYTES reserves K+1 per level)][(unused) pointer table][conv biases]
  // [down-sample biases][parked rows of streamed sequences]
  uint8_t* sc = reinterpret_cast<uint8_t*>(scratch);
  __nv_bfloat16* w_bf16 = reinterpret_cast<__nv_bfloat16*>(sc);
  const size_t w_bytes = (size_t)(1 + n_levels * (K + 1)) * kDim * kDim * 2;
  float* bias_dev = reinterpret_cast<float*>(sc + w_bytes + kK2PtrTableBytes);
  float* ds_bias_dev = bias_dev + HTCN_MAX_LEVELS * kDim;
  uint8_t* hist_dev = reinterpret_cast<uint8_t*>(ds_bias_dev + HTCN_MAX_LEVELS * kDim) + 256;   // per-chain parked rows
  k2_prepare_weights<<<n_wt + 1, 256, 0, st>>>(pa, w_bf16, bias_dev, ds_bias_dev);
  HTCN_LAUNCH_CHECK("k2_prepare_weights");
  CUtensorMap tw;
  int32_t rc = make_tmap_bf16(&tw, w_bf16, (uint64_t)n_wt * kDim, kDim, kDim, 64, 128, 128);
  if (rc) return rc;
  const size_t smem = sizeof(K2Smem) + 1024;
  // HTCN_K2_MULTICAST=1: CTA pairs with multicast weights.  Measured: 0.727 vs 0.719 ms (hier), 1.49 vs 1.55 ms (cfg3) --
  // halving the L2 -> SM weight traffic buys nothing, the kernel is bound by the latency of its 2-deep weight ring and of
  // the dependent layer chain, not by L2 bandwidth; independent CTAs stay the default.
  const char* mc = getenv("HTCN_K2_MULTICAST");
  const bool pair = mc && atoi(mc) != 0 && units >= 2;
  bool stream = false;
  for (int s = 0; s < slots.n; ++s) stream |= g.slot[s].chunks > 1;
  const bool aux = g.ds_mask != 0 || drop || h_save || a_save;
  // one CTA per SM running two tile chains over a 4-deep weight ring (k2_tcn_bf16_dual) for streamed long sequences, the
  // single-chain kernel (two CTAs per SM, 2-deep ring) otherwise -- measured: config 3 (4096 x 256, 4 levels) 0.817 vs
  // 0.856 ms, config 2 (40960 x 20, 2 levels) 0.638 vs 0.612 ms: both variants stream 352 KB of weights per tile out of L2
  // (~53 % of the chip's L2 throughput) and serialise a tile's MMA and epilogue phases, the deeper ring only helps the longer
  // layer chains.  HTCN_K2_DUAL=1 / 0 forces one or the other.
  const char* dual_env = getenv("HTCN_K2_DUAL");
  // The default for the plain inference stack: four tile chains per CTA sharing every weight tile (k2_tcn_quad.cu).
  // HTCN_K2_QUAD=0 (or any HTCN_K2_DUAL / HTCN_K2_MULTICAST choice) selects the kernels of this file.
  const char* quad_env = getenv("HTCN_K2_QUAD");
  if (!aux && !pair && !dual_env && (quad_env ? atoi(quad_env) != 0 : true))
    return k2_launch_quad(tw, g, (const __nv_bfloat16*)xe, sbias, (const float*)bias_dev, out_row, (__nv_bfloat16*)hout, hist_dev,
                          stream, st);
  // (with the issuer's descriptors advanced by their low words the two-chain kernel wins at both shapes: 0.523 vs 0.558 ms at
  // config 2, 0.718 vs 0.790 at config 3 -- it is the default for everything the four-chain kernel does not take)
  const bool dual = dual_env ? atoi(dual_env) != 0 : true;
  if (!pair && dual) {
    const size_t smem_d = sizeof(K2SmemDual) + 1024;
    const int grid_d = (units + 1) / 2 < 148 ? (units + 1) / 2 : 148;
#define HTCN_K2_LAUNCH_DUAL(STREAM, AUX)                                                                                 \
  do {                                                                                                                  \
    auto kern = k2_tcn_bf16_dual<STREAM, AUX>;                                                                          \
    HTCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));                    \
    kern<<<grid_d, kK2DualThreads, smem_d, st>>>(tw, g, (const __nv_bfloat16*)xe, sbias, (const float*)bias_dev,         \
                                                 (const float*)ds_bias_dev, out_row, (__nv_bfloat16*)hout,              \
                                                 (__nv_bfloat16*)h_save, (__nv_bfloat16*)a_save, hist_dev, drop);       \
    HTCN_LAUNCH_CHECK("k2_tcn_bf16_dual");                                                                              \
    return HTCN_OK;                                                                                                     \
  } while (0)
    if (stream && aux) HTCN_K2_LAUNCH_DUAL(true, true);
    if (stream) HTCN_K2_LAUNCH_DUAL(true, false);
    if (aux) HTCN_K2_LAUNCH_DUAL(false, true);
    HTCN_K2_LAUNCH_DUAL(false, false);
#undef HTCN_K2_LAUNCH_DUAL
  }
  const int grid = pair ? (units < 2 * 148 ? ((units + 1) & ~1) : 2 * 148) : (units < 2 * 148 ? units : 2 * 148);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(kK2Threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define HTCN_K2_LAUNCH(PAIR, STREAM, AUX)                                                                               \
  do {                                                                                                                  \
    auto kern = k2_tcn_bf16<PAIR, STREAM, AUX>;                                                                         \
    HTCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                      \
    HTCN_CUDA(cudaLaunchKernelEx(&cfg, kern, tw, g, (const __nv_bfloat16*)xe, sbias, (const float*)bias_dev,            \
                                 (const float*)ds_bias_dev, out_row, (__nv_bfloat16*)hout, (__nv_bfloat16*)h_save,      \
                                 (__nv_bfloat16*)a_save, hist_dev, drop));                                              \
    return HTCN_OK;                                                                                                     \
  } while (0)
  if (pair) {                      // multicast ablation: plain inference stacks only
    if (aux) {
      set_error("tcn_forward(bf16): HTCN_K2_MULTICAST supports the plain inference stack only");
      return HTCN_ERR_UNSUPPORTED;
    }
    if (stream) HTCN_K2_LAUNCH(true, true, false);
    HTCN_K2_LAUNCH(true, false, false);
  }
  if (stream && aux) HTCN_K2_LAUNCH(false, true, true);
  if (stream) HTCN_K2_LAUNCH(false, true, false);
  if (aux) HTCN_K2_LAUNCH(false, false, true);
  HTCN_K2_LAUNCH(false, false, false);
#undef HTCN_K2_LAUNCH
}

}  // namespace htcn
