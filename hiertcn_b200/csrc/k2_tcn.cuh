// Shared by the fused tcgen05 conv-stack kernels (k2_tcn_bf16.cu: one / two tile chains per CTA; k2_tcn_quad.cu: four tile
// chains per CTA sharing every weight tile): the tiling of the [B, T] positions into 128-row work units and the no-swizzle
// operand layout.  See k2_tcn_bf16.cu for the description of the layout.
#pragma once
#include "common.cuh"
#include "sm100.cuh"

namespace htcn {
using namespace sm100;

constexpr int kTR = 128;                 // rows per tile
constexpr int kMaxSpare = 32;            // supports (K-1)*d_max <= 32 rows of negative shift
constexpr int kRows = kTR + kMaxSpare;   // rows of the activation buffer (spare rows in front)
constexpr int kActBytes = 16 * kRows * 16;          // 16 channel chunks x rows x 16 B = 40 KB
constexpr int kWStageBytes = 2 * 128 * 128;         // one tap: [128 cout][128 cin] bf16, two 64-col swizzled chunks
constexpr int kWStages = 2;                         // 2 x 32 KB: with the 40 KB tile two CTAs fit one SM
constexpr int kK2EpiWarps = 8;           // 4 TMEM lane quarters x 2 channel halves: a thread owns 64 channels of one tile row
constexpr int kK2Threads = 32 * (kK2EpiWarps + 2);   // warps 0-7 epilogue/loader, warp 8 TMA producer, warp 9 MMA issuer
constexpr int kK2ProducerWarp = kK2EpiWarps, kK2MmaWarp = kK2EpiWarps + 1;

struct K2Slot {          // per session slot: tiling of its B sequences
  int off, L;            // first column in [B,T], length
  int unit0;             // first global work-unit index of this slot
  int seq_per_tile;      // > 0: short mode (unit = one tile of seq_per_tile sequences);  0: long mode (unit = one sequence)
  int chunks;            // tiles per unit: 1 (short), ceil(L / 128) (long)
};
struct K2Geom {
  int n_slots, n_units, B, T, K, n_levels;
  unsigned ds_mask;      // bit l: level l has a 1x1 down-sample residual (customized_tcn_cell.py:102-106): one more weight
                         // tile after the level's taps, accumulated into TMEM columns 128..255
  int P;                 // zero rows in front of each short sequence = max shift of the deepest level
  K2Slot slot[HTCN_MAX_SLOTS];
};
constexpr int kHistBytes = kMaxSpare * kDim * 2;      // one level's parked rows: [16 channel chunks][32 rows][16 B] = 8 KB

__device__ __forceinline__ int unit_slot(const K2Geom& g, int unit) {
  int s = 0;
  while (s + 1 < g.n_slots && g.slot[s + 1].unit0 <= unit) ++s;
  return s;
}
// tiles CTA `cta` of `n_cta` runs: the chunks of units cta, cta + n_cta, ...  (closed form per slot: the kernels call this
// once per thread and tile chain, a unit-by-unit walk cost ~25 000 clk of start-up at config 2)
__device__ __forceinline__ int cta_tile_count(const K2Geom& g, int cta, int n_cta) {
  int n = 0;
  for (int s = 0; s < g.n_slots; ++s) {
    const int a = g.slot[s].unit0, b = s + 1 < g.n_slots ? g.slot[s + 1].unit0 : g.n_units;
    int d = (cta - a) % n_cta;                      // first unit >= a congruent to cta
    if (d < 0) d += n_cta;
    const int first = a + d;
    if (first < b) n += ((b - first + n_cta - 1) / n_cta) * g.slot[s].chunks;
  }
  return n;
}

// no-swizzle K-major descriptor: core matrix = 8 rows x 16 B contiguous (SBO = 128 B), the two 16-byte K chunks of
// one K=16 step are LBO = kRows*16 B apart
__device__ __forceinline__ uint64_t make_desc_act(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((kRows * 16) >> 4) << 16;     // leading byte offset (K direction)
  d |= (uint64_t)(128 >> 4) << 32;              // stride byte offset (8-row groups)
  d |= (uint64_t)1 << 46;
  return d;                                     // layout type 0 = no swizzle
}

// row r of chunk `chunk` of work unit `unit` (unit >= n_units: a dummy tile, every row is padding)
__device__ __forceinline__ void tile_geometry(const K2Geom& g, int unit, int chunk, int r, const int* out_row, int& src,
                                              int& dst, int& sb, bool& own, int& slot) {
  src = -1; dst = -1; sb = 0; own = false; slot = 0;
  if (unit >= g.n_units) return;
  const int s = unit_slot(g, unit);
  slot = s;
  const K2Slot& sl = g.slot[s];
  const int lu = unit - sl.unit0;
  int b, t;
  bool is_out;
  if (sl.seq_per_tile > 0) {
    const int stride = sl.L + g.P;
    const int seg = r / stride;
    t = r % stride - g.P;
    b = lu * sl.seq_per_tile + seg;
    is_out = seg < sl.seq_per_tile;
  } else {
    b = lu;
    t = chunk * kTR + r;
    is_out = true;
  }
  const bool data = b < g.B && t >= 0 && t < sl.L;
  src = data ? b * g.T + sl.off + t : -1;
  sb = s * g.B + (b < g.B ? b : 0);
  own = data && is_out;            // this tile produces the row's values at every level
  if (own) dst = out_row ? out_row[src] : src;
}

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// weight tiles of one layer in the order the kernel consumes them: layer 0 = the in-projection, layer l >= 1 = the K taps of
// level l-1 followed by its down-sample kernel if it has one
__device__ __forceinline__ void k2_layer_tiles(const K2Geom& g, int layer, int& first, int& taps, bool& ds) {
  if (layer == 0) {
    first = 0; taps = 1; ds = false;
    return;
  }
  const int l = layer - 1;
  first = 1 + l * g.K + __popc(g.ds_mask & ((1u << l) - 1u));
  taps = g.K;
  ds = (g.ds_mask >> l) & 1u;
}

// k2_tcn_quad.cu
int32_t k2_launch_quad(const CUtensorMap& tw, const K2Geom& g, const __nv_bfloat16* xe, const float* sbias, const float* bias_dev,
                       const int* out_row, __nv_bfloat16* hout, uint8_t* hist_dev, bool stream, cudaStream_t st);

}  // namespace htcn
