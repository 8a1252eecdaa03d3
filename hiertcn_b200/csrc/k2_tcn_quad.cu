// K2 (bf16 tier), weight-sharing form: FOUR tile chains per CTA, every weight tile fetched ONCE for all four.
// Same arithmetic, layout and tiling as k2_tcn_bf16.cu (model_tcn.py:35 + customized_tcn_cell.py:46-49,109-127,147-161;
// bit-identical results); what changes is who pays for the weights.  The one- and two-chain kernels stream all weight tiles
// of the stack (352 KB at config 2, 672 KB at config 3) out of L2 for EVERY 128-row tile: at the measured ~34 B/clk of L2 -> SM
// ingress that is 10 000 clk per tile against 5 600 clk of MMAs, and it stretches every layer's MMA phase 2.7x
// (profiles/r2_k2_ncu.txt: tensor pipe active 39-41 %).  Here one CTA per SM owns four activation tiles (4 x 34..40 KB of
// shared memory, 4 x 128 TMEM columns = all of tensor memory) that walk the layers in lock step:
//   * the producer warp fetches each weight half-tap ([128 cout][64 cin] bf16 = 16 KB, one TMA box) once per ROUND
//     (= one layer of all four chains) into a 4..5-stage ring: 16 KB feed 4 chains x 4 MMAs = 1 024 clk of tensor work,
//     i.e. 16 B/clk -- under the ingress limit, so the ring no longer stalls the issuer;
//   * the MMA warp runs tap-major, chain-minor: for each half-tap, the 4 MMAs of chain 0, 1, 2, 3; a chain's accumulator is
//     committed right after its last MMA of the layer, so chain 0's epilogue starts while chains 1-3 still use the tensor
//     pipe, and the next layer's first tap starts with chain 0, whose operand is rewritten first;
//   * 16 epilogue warps, 4 per chain (one per TMEM lane quarter = one per scheduler): a thread owns one tile row with all
//     128 channels (bias / relu / residual / relu, bf16 rounding, in-place rewrite of the operand; causal pad rows stay zero).
// Chain c of CTA b is "virtual CTA" c * gridDim.x + b of the one-chain kernel (units, streaming hand-over scratch): small
// problems spread over the SMs first and only then stack chains on an SM.
// The plain inference stack only (no down-sample level -- that needs a second accumulator per chain --, no training
// dropout, no saved activations): those run on k2_tcn_bf16.cu's kernels.
#include <cstdlib>

#include "k2_tcn.cuh"

namespace htcn {
using namespace sm100;

constexpr int kQChains = 4;
constexpr int kQEpiWarps = 4 * kQChains;                 // warp w: chain w >> 2, TMEM lane quarter w & 3
// + kIssuers TMA producer warps + kIssuers MMA issuer warps (issuer w drives the chain groups [w, w+1) * kGroups / kIssuers
// through its own part of the weight ring)
constexpr int q_threads(int issuers) { return 32 * (kQEpiWarps + 2 * issuers); }
constexpr int kQStageBytes = 128 * 64 * 2;               // half a tap: [128 cout][64 cin] bf16, 128-byte swizzle

// kSpare = rows in front of a tile that a shifted tap may read (>= (K-1)*d_max; 32 when long sequences are streamed: the
// parked rows of the previous chunk land there).  Fewer spare rows = a smaller operand tile = a deeper weight ring.
template <int kSpare, bool kFullTap = false>
struct QuadCfg {
  static constexpr int kRowsQ = kTR + kSpare;
  static constexpr int kActQ = 16 * kRowsQ * 16;          // 16 channel chunks x rows x 16 B
  // ring stages: half-taps (16 KB, 4..5 of them) or -- kFullTap -- two whole taps (32 KB: half the waits / commits / stage
  // bookkeeping per MMA for the issuer)
  static constexpr int kHalves = kFullTap ? 2 : 1;
  static constexpr int kStageB = kQStageBytes * kHalves;
  static constexpr int kStages = kFullTap ? 2 : (kSpare == 8 ? 5 : 4);
  static constexpr bool kBiasSmem = kSpare < 32;          // with 4 x 40 KB tiles the level biases stay in global memory (L1)
  struct alignas(1024) Smem {
    uint8_t w[kStages][kStageB];
    uint8_t act[kQChains][kActQ];
    float bias[kBiasSmem ? HTCN_MAX_LEVELS : 1][kDim];
    uint64_t w_full[kStages], w_empty[kStages], acc_ready[kQChains], act_ready[kQChains];
    uint32_t tmem_base;
  };
  static_assert(sizeof(Smem) + 1024 <= 232448, "shared-memory budget of one CTA");
};

// no-swizzle K-major descriptor of an operand tile with kRowsT rows (see make_desc_act)
template <int kRowsT>
__device__ __forceinline__ uint64_t make_desc_act_t(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((kRowsT * 16) >> 4) << 16;    // leading byte offset (K direction)
  d |= (uint64_t)(128 >> 4) << 32;              // stride byte offset (8-row groups)
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// relu of two packed bf16 values
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t x) {
  uint32_t y;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(y) : "r"(x), "r"(0u));
  return y;
}

template <int kSpare, bool kStream, int kGroups, int kIssuers, bool kFullTap>
__global__ void __launch_bounds__(q_threads(kIssuers), 1)
k2_tcn_bf16_quad(const __grid_constant__ CUtensorMap tmap_w, K2Geom g, const __nv_bfloat16* __restrict__ xe,
                 const float* __restrict__ sbias, const float* __restrict__ bias_all /*[n_levels][128]*/,
                 const int* __restrict__ out_row, __nv_bfloat16* __restrict__ hout,
                 uint8_t* __restrict__ hist /* [4 * gridDim.x][2][n_levels][kHistBytes] parked level inputs of streamed sequences */,
                 int kLag /* rounds group g runs behind group g-1 */) {
  using C = QuadCfg<kSpare, kFullTap>;
  using Smem = typename C::Smem;
  static_assert(!kStream || kSpare == kMaxSpare, "streamed chunks hand over kMaxSpare rows");
  constexpr int kRowsQ = C::kRowsQ;
  constexpr int kPerGroup = kQChains / kGroups;             // chains that share one fetch of a layer's weights
  constexpr int kQThreads = q_threads(kIssuers);
  constexpr int kGroupsPerIssuer = kGroups / kIssuers;
  static_assert(kGroups % kIssuers == 0 && C::kStages >= 2 * kIssuers, "issuer split");
  // ring of issuer w: stages [ring0, ring0 + ring_n) -- the first issuer takes the odd stage
  constexpr int kRing0N = kIssuers == 1 ? C::kStages : (C::kStages + 1) / 2;
  // Group g runs kLag rounds behind group g-1, so in any round the groups are at DIFFERENT layers: the short in-projection
  // round (1 tap) of one group sits beside conv rounds (K taps) of the others instead of all chains idling the tensor pipe
  // through the same short round and the input staging of a new tile
  extern __shared__ uint8_t smem_raw[];
  auto& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_layers = g.n_levels + 1;                       // layer 0 = in-projection
  const int n_vc = kQChains * (int)gridDim.x;                // virtual CTAs = tile chains
  int steps[kQChains];                                       // layers each chain runs
  int n_rounds = 0;
#pragma unroll
  for (int c = 0; c < kQChains; ++c) {
    steps[c] = cta_tile_count(g, c * (int)gridDim.x + (int)blockIdx.x, n_vc) * n_layers;
    const int end = steps[c] + (c / kPerGroup) * kLag;
    n_rounds = end > n_rounds ? end : n_rounds;
  }

  if (tid == 0) {
    prefetch_tmap(&tmap_w);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&sm.w_full[s], 1);
      mbar_init(&sm.w_empty[s], 1);
    }
    for (int c = 0; c < kQChains; ++c) {
      mbar_init(&sm.acc_ready[c], 1);
      mbar_init(&sm.act_ready[c], 128);                      // every row owner of the chain
    }
    fence_barrier_init();
  }
  if (C::kBiasSmem)
    for (int i = tid; i < g.n_levels * kDim; i += kQThreads) sm.bias[i / kDim][i % kDim] = bias_all[i];
  for (int i = tid; i < kQChains * C::kActQ / 16; i += kQThreads)
    reinterpret_cast<uint4*>(&sm.act[0][0])[i] = make_uint4(0, 0, 0, 0);
  if (warp == kQEpiWarps + kIssuers) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, sm.tmem_base, 0);

  if (warp >= kQEpiWarps && warp < kQEpiWarps + kIssuers) {
    // ===================== weight producer: the half-taps of layer (i mod n_layers), once per round =====================
    const int iw = warp - kQEpiWarps;
    const uint32_t ring0 = iw == 0 ? 0u : (uint32_t)kRing0N, ring_n = iw == 0 ? (uint32_t)kRing0N : (uint32_t)(C::kStages - kRing0N);
    if (lane == 0) {
      uint32_t stage = ring0, wphase = 0;
      for (int i = 0; i < n_rounds; ++i) {
#pragma unroll
        for (int grp = 0; grp < kGroups; ++grp) {
          if (grp / kGroupsPerIssuer != iw) continue;
          const int il = i - grp * kLag;                         // the group's own round
          bool active = false;
#pragma unroll
          for (int c = grp * kPerGroup; c < (grp + 1) * kPerGroup; ++c) active |= il >= 0 && il < steps[c];
          if (!active) continue;
          int first, taps;
          bool ds;
          k2_layer_tiles(g, il % n_layers, first, taps, ds);
          for (int j = first; j < first + taps; ++j) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              if (half % C::kHalves == 0) {
                mbar_wait_relaxed(&sm.w_empty[stage], wphase ^ 1);
                mbar_arrive_expect_tx(&sm.w_full[stage], C::kStageB);
              }
              tma_load_2d(sm.w[stage] + (half % C::kHalves) * kQStageBytes, &tmap_w, 64 * half, j * 128, &sm.w_full[stage]);
              if (half % C::kHalves == C::kHalves - 1 && ++stage == ring0 + ring_n) {
                stage = ring0;
                wphase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp >= kQEpiWarps + kIssuers) {
    // ===================== MMA issuer(s): tap-major, chain-minor =====================
    // (the whole warp runs the uniform waits and descriptor arithmetic, one elected lane issues: see k2_tcn_bf16.cu)
    const int iw = warp - kQEpiWarps - kIssuers;
    const uint32_t ring0 = iw == 0 ? 0u : (uint32_t)kRing0N, ring_n = iw == 0 ? (uint32_t)kRing0N : (uint32_t)(C::kStages - kRing0N);
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(kTR, 128);
    // descriptor halves (make_desc_act / make_desc_k_sw128): the issuer only ever adds 16-byte units to the low words
    constexpr uint32_t kAHi = (128u >> 4) | (1u << 14), kBHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = (((smem_u32(&sm.act[0][0]) + kSpare * 16) & 0x3FFFF) >> 4) | ((uint32_t)kRowsQ << 16);
    const uint32_t b_lo0 = ((smem_u32(&sm.w[0][0]) & 0x3FFFF) >> 4) | (1u << 16);
    uint32_t stage = ring0, wphase = 0;
    for (int i = 0; i < n_rounds; ++i) {
#pragma unroll
      for (int grp = 0; grp < kGroups; ++grp) {
        if (grp / kGroupsPerIssuer != iw) continue;
        const int il = i - grp * kLag;                           // the group's own round
        bool active = false;
#pragma unroll
        for (int c = grp * kPerGroup; c < (grp + 1) * kPerGroup; ++c) active |= il >= 0 && il < steps[c];
        if (!active) continue;
        const int layer = il % n_layers;
        const int taps = layer == 0 ? 1 : g.K;
        const int dil = layer == 0 ? 1 : (1 << (layer - 1));
        for (int tap = 0; tap < taps; ++tap) {
          const uint32_t shift = (uint32_t)((taps - 1 - tap) * dil);   // rows back in time (customized_tcn_cell.py:46-49)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (half % C::kHalves == 0) {
              mbar_wait(&sm.w_full[stage], wphase);
              tc_fence_after_sync();
            }
            const uint32_t b_lo = b_lo0 + stage * (C::kStageB >> 4) + (half % C::kHalves) * (kQStageBytes >> 4);
#pragma unroll
            for (int c = grp * kPerGroup; c < (grp + 1) * kPerGroup; ++c) {
              if (il >= steps[c]) continue;
              if (tap == 0 && half == 0) {
                mbar_wait(&sm.act_ready[c], (uint32_t)(il & 1)); // the chain's operand tile is written + fenced
                tc_fence_after_sync();
              }
              // tap = the same tile read `shift` rows (16 B each) further back; K chunk pair k = 2k * kRowsQ rows further on
              const uint32_t a_lo = a_lo0 + (uint32_t)c * (C::kActQ >> 4) - shift + (uint32_t)(half * 8) * kRowsQ;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                if (leader)
                  umma_bf16_lohi(tmem + (uint32_t)c * 128, a_lo + (uint32_t)(2 * kk) * kRowsQ, kAHi, b_lo + kk * 2, kBHi, idesc,
                                 (tap | half | kk) != 0);
              if (tap == taps - 1 && half == 1 && leader) umma_commit(&sm.acc_ready[c]);
            }
            if (half % C::kHalves == C::kHalves - 1) {
              if (leader) umma_commit(&sm.w_empty[stage]);
              if (++stage == ring0 + ring_n) {
                stage = ring0;
                wphase ^= 1;
              }
            }
          }
        }
      }
    }
  } else {
    // ===================== loader + epilogue: thread = one tile row of its chain, all 128 channels =====================
    const int cn = warp >> 2;                                   // tile chain of this worker warp
    const int vc = cn * (int)gridDim.x + (int)blockIdx.x;       // its virtual CTA
    const int r = tid & 127;                                    // TMEM lane r (warp w may touch lanes 32(w%4)..+31)
    const uint32_t tmem_c = tmem + (uint32_t)cn * 128 + ((uint32_t)((warp & 3) * 32) << 16);
    uint8_t* act_c = sm.act[cn];
    const uint32_t my_act = smem_u32(act_c) + (kSpare + r) * 16;      // + c * kRowsQ * 16 for channel chunk c
    const int my_tiles = cta_tile_count(g, vc, n_vc);
    long long n_acc = 0;
    int unit = vc, chunk = 0, unit_chunks = unit < g.n_units ? g.slot[unit_slot(g, unit)].chunks : 1;
    // streaming of long sequences: the threads of the tile's last kMaxSpare rows own the hand-over to the next chunk
    const bool hist_owner = kStream && r >= kTR - kMaxSpare;
    const int hj = r - (kTR - kMaxSpare);                       // spare row / parked row of this thread
    uint8_t* my_hist = hist + (size_t)vc * 2 * g.n_levels * kHistBytes;
    bool spare_dirty = false;                                   // the spare rows hold parked data (not the zero pad)
    for (int it = 0; it < my_tiles; ++it) {
      int src, dst, sb, slot_idx;
      bool own;
      tile_geometry(g, unit, chunk, r, out_row, src, dst, sb, own, slot_idx);
      const bool streaming = kStream && unit_chunks > 1;
      // the tile after this one: its input rows are fetched under this tile's last epilogue (see below)
      int n_unit = unit, n_chunk = chunk + 1, n_unit_chunks = unit_chunks;
      if (n_chunk == unit_chunks) {
        n_unit = unit + n_vc;
        n_chunk = 0;
        n_unit_chunks = n_unit < g.n_units ? g.slot[unit_slot(g, n_unit)].chunks : 1;
      }
      const bool has_next = it + 1 < my_tiles;
      {
        // ---- stage the input row (bf16 Xe) into the operand layout; zero rows stay zero.  (Fetching the next tile's rows
        // with cp.async under the previous tile's last epilogue was tried: 32 lanes x 16 B from 32 different global rows
        // arrive one by one and cost ~20 shared-memory write wavefronts per instruction instead of 4 -- slower.  The rows are
        // prefetched into L2 there instead.)
        const uint4* p = src >= 0 ? reinterpret_cast<const uint4*>(xe + (long long)src * kDim) : nullptr;
        uint4 xv[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) xv[c] = p ? __ldg(p + c) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int c = 0; c < 16; ++c) sts_u4(my_act + c * (kRowsQ * 16), xv[c]);
        fence_proxy_async_smem();
        mbar_arrive(&sm.act_ready[cn]);
      }
      for (int layer = 0; layer < n_layers; ++layer, ++n_acc) {
        const bool last = layer == n_layers - 1;
        // the per-(slot, user) bias of the in-projection (model_hier.py:91 hoisted: state . W_in[D:]): 32 channels of it in
        // flight before the accumulator is waited for
        const float* sb_row = (layer == 0 && sbias && src >= 0) ? sbias + (long long)sb * kDim : nullptr;
        float4 sba[4], sbb[4];                                   // 16 channels each, one piece ahead of the math
        if (layer == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) sba[i] = sb_row ? __ldg(reinterpret_cast<const float4*>(sb_row) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(&sm.acc_ready[cn], (uint32_t)(n_acc & 1));
        tc_fence_after_sync();
        // The MMAs that read the spare rows have retired: refill them for the NEXT layer (conv level `layer`), whose
        // taps reach back up to kMaxSpare rows -- with the rows chunk-1 parked for that level, or with the causal zero pad.
        uint8_t* park = my_hist + ((size_t)(chunk & 1) * g.n_levels + layer) * kHistBytes;
        if (kStream && !last && hist_owner) {
          const uint8_t* prev = my_hist + ((size_t)((chunk & 1) ^ 1) * g.n_levels + layer) * kHistBytes;
          if (streaming && chunk > 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c) cp_async_16(act_c + c * (kRowsQ * 16) + hj * 16, prev + (c * kMaxSpare + hj) * 16);
            spare_dirty = true;
          } else if (spare_dirty) {                              // back to the causal zero pad
#pragma unroll
            for (int c = 0; c < 16; ++c) sts_u4(smem_u32(act_c) + c * (kRowsQ * 16) + hj * 16, make_uint4(0, 0, 0, 0));
            spare_dirty = false;
          }
        }
        // park this row of the next level's input for the sequence's next chunk (read back by this same thread)
        const bool do_park = kStream && !last && streaming && hist_owner && chunk + 1 < unit_chunks;
        if (layer == 0) {
          // ---- in-projection: acc + sbias -> bf16, in 16-column pieces; the next piece's accumulator columns and bias values
          // are in flight under the math of the current one
          auto piece0 = [&](const uint32_t (&v)[16], const float4 (&sbv)[4], int c0) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const float2 a0 = fadd2(make_float2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1])), make_float2(sbv[2 * q].x, sbv[2 * q].y));
              const float2 a1 = fadd2(make_float2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3])), make_float2(sbv[2 * q].z, sbv[2 * q].w));
              const float2 a2 = fadd2(make_float2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5])), make_float2(sbv[2 * q + 1].x, sbv[2 * q + 1].y));
              const float2 a3 = fadd2(make_float2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7])), make_float2(sbv[2 * q + 1].z, sbv[2 * q + 1].w));
              const uint4 packed = make_uint4(pack_bf16x2(a0.x, a0.y), pack_bf16x2(a1.x, a1.y), pack_bf16x2(a2.x, a2.y), pack_bf16x2(a3.x, a3.y));
              const int c = c0 + q;
              if (!last) {
#ifndef HTCN_K2Q_NOSMEM   /* timing experiment only (-DHTCN_K2Q_NOSMEM): the epilogue without its shared-memory traffic */
                if (src >= 0) sts_u4(my_act + c * (kRowsQ * 16), packed);          // causal pad rows are never written: they stay zero
#else
                if (src == -12345) sts_u4(my_act + c * (kRowsQ * 16), packed);
#endif
                if (do_park) *reinterpret_cast<uint4*>(park + (c * kMaxSpare + hj) * 16) = src >= 0 ? packed : make_uint4(0, 0, 0, 0);
              } else if (dst >= 0) reinterpret_cast<uint4*>(hout + (long long)dst * kDim)[c] = packed;
            }
          };
          auto sb_load = [&](float4 (&sbv)[4], int pc) {
#pragma unroll
            for (int i = 0; i < 4; ++i) sbv[i] = sb_row ? __ldg(reinterpret_cast<const float4*>(sb_row + pc * 16) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          };
          uint32_t va[16], vb[16];
          tmem_ld_32x16(tmem_c, va);
          tmem_ld_wait(va);
#pragma unroll
          for (int pc = 0; pc < 8; pc += 2) {
            tmem_ld_32x16(tmem_c + (pc + 1) * 16, vb);
            sb_load(sbb, pc + 1);
            piece0(va, sba, pc * 2);
            tmem_ld_wait(vb);
            if (pc + 2 < 8) {
              tmem_ld_32x16(tmem_c + (pc + 2) * 16, va);
              sb_load(sba, pc + 2);
            }
            piece0(vb, sbb, pc * 2 + 2);
            if (pc + 2 < 8) tmem_ld_wait(va);
          }
        }
        if (layer > 0) {
          // ---- conv level: relu(relu(acc + b) + residual) -> bf16 (customized_tcn_cell.py:109-127).  The accumulator is read
          // in 16-column pieces, the next piece in flight under the math of the current one; adds are packed (add.f32x2), the
          // outer relu acts on the rounded pair (max.bf16x2: rounding is monotonic and keeps the sign, so relu and rounding
          // commute).  The residual loads run one piece ahead too; the bias table is read with movable (non-volatile) loads.
          const float* bias_g = bias_all + (layer - 1) * kDim;
          const uint32_t bias_s = smem_u32(sm.bias[C::kBiasSmem ? layer - 1 : 0]);
          if (last && has_next) {                              // the next tile's input row: on its way into L2
            int nsrc, ndst, nsb, nslot;
            bool nown;
            tile_geometry(g, n_unit, n_chunk, r, out_row, nsrc, ndst, nsb, nown, nslot);
            if (nsrc >= 0) {
              prefetch_l2(xe + (long long)nsrc * kDim);
              prefetch_l2(xe + (long long)nsrc * kDim + 64);
            }
          }
          // residual = this row's input to the level (bf16 x 8 per 16-byte slot)
          auto res_load = [&](uint4 (&res)[2], int c0) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const uint32_t slot = my_act + (c0 + q) * (kRowsQ * 16);
#ifndef HTCN_K2Q_NOSMEM
              res[q] = lds_u4(slot);
#else
              res[q] = make_uint4(slot, slot, slot, slot);
#endif
            }
          };
          auto piece = [&](const uint32_t (&v)[16], const uint4 (&res)[2], int c0) {    // channels 8*c0 .. 8*c0+15 of this row
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int c = c0 + q;
              float4 bv0, bv1;
              if (C::kBiasSmem) {
                bv0 = lds_f4_const(bias_s + c * 32);
                bv1 = lds_f4_const(bias_s + c * 32 + 16);
              } else {
                bv0 = __ldg(reinterpret_cast<const float4*>(bias_g + c * 8));
                bv1 = __ldg(reinterpret_cast<const float4*>(bias_g + c * 8) + 1);
              }
              const float2 b2[4] = {make_float2(bv0.x, bv0.y), make_float2(bv0.z, bv0.w), make_float2(bv1.x, bv1.y),
                                    make_float2(bv1.z, bv1.w)};
              const uint32_t rw[4] = {res[q].x, res[q].y, res[q].z, res[q].w};
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 a = fadd2(make_float2(__uint_as_float(v[q * 8 + 2 * e]), __uint_as_float(v[q * 8 + 2 * e + 1])), b2[e]);
                a.x = fmaxf(a.x, 0.f);                                                           // relu(conv + b)
                a.y = fmaxf(a.y, 0.f);
                a = fadd2(a, make_float2(bf16_lo(rw[e]), bf16_hi(rw[e])));                       // + residual
                pk[e] = relu_bf16x2(pack_bf16x2(a.x, a.y));                                      // relu, on the rounded pair
              }
              const uint4 packed = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              if (!last) {
#ifndef HTCN_K2Q_NOSMEM   /* timing experiment only (-DHTCN_K2Q_NOSMEM): the epilogue without its shared-memory traffic */
                if (src >= 0) sts_u4(my_act + c * (kRowsQ * 16), packed);          // causal pad rows are never written: they stay zero
#else
                if (src == -12345) sts_u4(my_act + c * (kRowsQ * 16), packed);
#endif
                if (do_park) *reinterpret_cast<uint4*>(park + (c * kMaxSpare + hj) * 16) = src >= 0 ? packed : make_uint4(0, 0, 0, 0);
              } else if (dst >= 0) reinterpret_cast<uint4*>(hout + (long long)dst * kDim)[c] = packed;
            }
          };
          uint32_t va[16], vb[16];
          uint4 ra[2], rb[2];
          tmem_ld_32x16(tmem_c, va);
          res_load(ra, 0);
          tmem_ld_wait(va);
#pragma unroll
          for (int pc = 0; pc < 8; pc += 2) {                    // 8 pieces of 16 columns, two per trip
            tmem_ld_32x16(tmem_c + (pc + 1) * 16, vb);
            res_load(rb, pc * 2 + 2);
            piece(va, ra, pc * 2);
            tmem_ld_wait(vb);
            if (pc + 2 < 8) {
              tmem_ld_32x16(tmem_c + (pc + 2) * 16, va);
              res_load(ra, pc * 2 + 4);
            }
            piece(vb, rb, pc * 2 + 2);
            if (pc + 2 < 8) tmem_ld_wait(va);
          }
        }
        tc_fence_before_sync();
        if (!last) {
          if (kStream && hist_owner) cp_async_wait_all();        // the parked rows have landed in the spare rows
          fence_proxy_async_smem();
          mbar_arrive(&sm.act_ready[cn]);
        }
      }
      unit = n_unit;
      chunk = n_chunk;
      unit_chunks = n_unit_chunks;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kQEpiWarps + kIssuers) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem);
  }
}

template <int kSpare, bool kStream, int kGroups, int kIssuers, bool kFullTap = false>
static int32_t launch_quad_g(const CUtensorMap& tw, const K2Geom& g, const __nv_bfloat16* xe, const float* sbias,
                           const float* bias_dev, const int* out_row, __nv_bfloat16* hout, uint8_t* hist_dev, cudaStream_t st) {
  const size_t smem = sizeof(typename QuadCfg<kSpare, kFullTap>::Smem) + 1024;
  const int grid = g.n_units < 148 ? g.n_units : 148;
  auto kern = k2_tcn_bf16_quad<kSpare, kStream, kGroups, kIssuers, kFullTap>;
  HTCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const char* lag_env = getenv("HTCN_K2_LAG");
  kern<<<grid, q_threads(kIssuers), smem, st>>>(tw, g, xe, sbias, bias_dev, out_row, hout, hist_dev, lag_env ? atoi(lag_env) : 1);
  HTCN_LAUNCH_CHECK("k2_tcn_bf16_quad");
  return HTCN_OK;
}

template <int kSpare, bool kStream>
static int32_t launch_quad(const CUtensorMap& tw, const K2Geom& g, const __nv_bfloat16* xe, const float* sbias,
                           const float* bias_dev, const int* out_row, __nv_bfloat16* hout, uint8_t* hist_dev, cudaStream_t st) {
  // HTCN_K2_QUAD = number of chain groups: 1 = all four chains in lock step (every weight tile fetched once per round, but
  // the MMA and epilogue phases of a round do not overlap), 2 = two pairs (each pair fetches its own copy of the layer's
  // weights), 4 = four independent chains; group g runs HTCN_K2_LAG (default 1) rounds behind group g-1.  Measured
  // (profiles/r2_k2_quad_sweep.txt): short sequences (config 2) 0.476 (4 groups, no lag) -> 0.436 ms (4 groups, lag 1), pairs
  // 0.467; streamed long sequences (config 3) within 2 % of each other -- pairs (half the weight traffic) are the default there.
  const char* e = getenv("HTCN_K2_QUAD");
  const int groups = e ? atoi(e) : (kStream ? 2 : 4);
  // HTCN_K2_ISSUERS = 2: a second MMA-issuing warp (and producer warp): the chains 0-1 and 2-3 get an issuer and a part of
  // the weight ring each -- with four independent chains ONE issuer spends ~59 instructions per half-tap of one chain, about the
  // 256 clk of its 4 MMAs
  const char* ie = getenv("HTCN_K2_ISSUERS");
  const int issuers = ie ? atoi(ie) : 1;
  // HTCN_K2_FULLTAP: chain pairs / lock step over two whole-tap stages instead of 4-5 half-tap stages (half the waits, commits
  // and stage bookkeeping per MMA in the issuer): pairs 0.476 -> 0.457 ms at config 2, 0.730 -> 0.711 ms at config 3 -- the
  // default for streamed sequences (where pairs are the default)
  const char* fe = getenv("HTCN_K2_FULLTAP");
  const bool fulltap = fe ? atoi(fe) != 0 : kStream;
  if (fulltap && groups == 2) return launch_quad_g<kSpare, kStream, 2, 1, true>(tw, g, xe, sbias, bias_dev, out_row, hout, hist_dev, st);
  if (fulltap && groups == 1) return launch_quad_g<kSpare, kStream, 1, 1, true>(tw, g, xe, sbias, bias_dev, out_row, hout, hist_dev, st);
  if (issuers == 2 && groups == 4) return launch_quad_g<kSpare, kStream, 4, 2>(tw, g, xe, sbias, bias_dev, out_row, hout, hist_dev, st);
  if (issuers == 2 && groups == 2) return launch_quad_g<kSpare, kStream, 2, 2>(tw, g, xe, sbias, bias_dev, out_row, hout, hist_dev, st);
  if (groups == 1) return launch_quad_g<kSpare, kStream, 1, 1>(tw, g, xe, sbias, bias_dev, out_row, hout, hist_dev, st);
  if (groups == 4) return launch_quad_g<kSpare, kStream, 4, 1>(tw, g, xe, sbias, bias_dev, out_row, hout, hist_dev, st);
  return launch_quad_g<kSpare, kStream, 2, 1>(tw, g, xe, sbias, bias_dev, out_row, hout, hist_dev, st);
}

int32_t k2_launch_quad(const CUtensorMap& tw, const K2Geom& g, const __nv_bfloat16* xe, const float* sbias, const float* bias_dev,
                       const int* out_row, __nv_bfloat16* hout, uint8_t* hist_dev, bool stream, cudaStream_t st) {
  if (stream) return launch_quad<32, true>(tw, g, xe, sbias, bias_dev, out_row, hout, hist_dev, st);
  if (g.P <= 8) return launch_quad<8, false>(tw, g, xe, sbias, bias_dev, out_row, hout, hist_dev, st);
  if (g.P <= 16) return launch_quad<16, false>(tw, g, xe, sbias, bias_dev, out_row, hout, hist_dev, st);
  return launch_quad<32, false>(tw, g, xe, sbias, bias_dev, out_row, hout, hist_dev, st);
}

}  // namespace htcn
