// Weight gradients of the conv stack on the 5th-gen tensor cores (bf16 tier of the training step).
//
//     dW_l[tap][cin][cout] += sum_r h_l[src(r, tap)][cin] * dp[r][cout],   src(r, tap) = r - (K-1-tap) d, zero before the
//                                                                          start of r's sequence (customized_tcn_cell.py:46-49)
// is a "TN" GEMM whose contraction runs over the B*T positions (819 200 at config 5) with M = N = 128: 2.1 GFLOP per tap
// that the fp32 split-K kernel (sgemm_tn_atomic) runs at FFMA speed -- 8.3 ms of the 49 ms step.  Here both operands are
// first written TRANSPOSED and ZERO-PADDED as bf16 [128][Kp] (pad_transpose_bf16: every sequence is preceded by P = the
// largest shift zero columns, so a shifted read can never reach the previous sequence -- the same physical-padding idea as
// the forward conv kernel).  The tap shift runs along the CONTRACTION (innermost, contiguous) dimension, and a TMA box must
// start 16-byte aligned there (a box at column k0 - shift faults as an illegal instruction for odd shifts), so the A
// operand is written once PER TAP, already shifted (pad_transpose reads a 64-column tile plus the largest shift once and
// writes every tap's copy).  Every tap is then a plain K-major x K-major tcgen05 GEMM.  One CTA accumulates up to 3 taps
// (3 x 128 TMEM columns) over its share of the columns and adds its partial to dW with float4 atomics.
#include "train.cuh"
#include "sm100.cuh"

namespace htcn {
using namespace sm100;

namespace {

constexpr int kWgStages = 3;
constexpr int kWgTile = 128 * 128;                // one [128 rows x 64 columns] bf16 tile, 128B swizzle
constexpr int kWgMaxTaps = 3;                     // per launch: 3 x 128 accumulator columns (TMEM holds 512)
constexpr int kWgThreads = 192;                   // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue

struct alignas(1024) WgSmem {
  uint8_t a[kWgStages][kWgMaxTaps][kWgTile];
  uint8_t b[kWgStages][kWgTile];
  uint64_t full[kWgStages], empty[kWgStages], done;
  uint32_t tmem_base;
};

struct WgTaps {
  int n;
};

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
                  const __grid_constant__ CUtensorMap tmap_a2, const __grid_constant__ CUtensorMap tmap_b, int n_steps,
                  WgTaps taps, float* __restrict__ dW /* [taps.n][128][128], accumulated */) {
  extern __shared__ uint8_t smem_raw[];
  auto& sm = *reinterpret_cast<WgSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s0 = (int)((long long)n_steps * blockIdx.x / gridDim.x);
  const int s1 = (int)((long long)n_steps * (blockIdx.x + 1) / gridDim.x);
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a0);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, sm.tmem_base, 0);

  if (warp == 0) {
    if (lane == 0) {
      for (int i = s0; i < s1; ++i) {
        const int n = i - s0, s = n % kWgStages;
        mbar_wait_relaxed(&sm.empty[s], (uint32_t)(((n / kWgStages) & 1) ^ 1));
        mbar_arrive_expect_tx(&sm.full[s], (uint32_t)((1 + taps.n) * kWgTile));
        const int k0 = i * 64;
        tma_load_2d(sm.b[s], &tmap_b, k0, 0, &sm.full[s]);
        tma_load_2d(sm.a[s][0], &tmap_a0, k0, 0, &sm.full[s]);
        if (taps.n > 1) tma_load_2d(sm.a[s][1], &tmap_a1, k0, 0, &sm.full[s]);
        if (taps.n > 2) tma_load_2d(sm.a[s][2], &tmap_a2, k0, 0, &sm.full[s]);
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(128, 128);
    for (int i = s0; i < s1; ++i) {
      const int n = i - s0, s = n % kWgStages;
      mbar_wait(&sm.full[s], (uint32_t)((n / kWgStages) & 1));
      tc_fence_after_sync();
      for (int t = 0; t < taps.n; ++t) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = make_desc_k_sw128(smem_u32(sm.a[s][t]) + k * 32);
          const uint64_t db = make_desc_k_sw128(smem_u32(sm.b[s]) + k * 32);
          if (leader) umma_bf16(tmem + t * 128, da, db, idesc, (n | k) != 0);
        }
      }
      if (leader) umma_commit(&sm.empty[s]);
    }
    if (leader) umma_commit(&sm.done);
  } else if (s1 > s0) {
    // epilogue: TMEM lane = cin row; each warp owns the lane quarter warp % 4
    mbar_wait(&sm.done, 0);
    tc_fence_after_sync();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    for (int t = 0; t < taps.n; ++t) {
      float* dst = dW + ((long long)t * 128 + row) * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + t * 128 + c * 32, v);
        tmem_ld_wait(v);
#pragma unroll
        for (int u = 0; u < 32; u += 4)
          atomicAdd(reinterpret_cast<float4*>(dst + c * 32 + u),
                    make_float4(__uint_as_float(v[u]), __uint_as_float(v[u + 1]), __uint_as_float(v[u + 2]), __uint_as_float(v[u + 3])));
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem);
  }
}

// ---- the same product with MN-MAJOR operands: no transposes -------------------------------------------------------------
// dW[t][cin][cout] += sum_r hP[r - shift_t][cin] dpP[r][cout] over zero-padded ROW-major bf16 copies hP, dpP [Kp][128] (every
// sequence preceded by P zero rows).  Both operands are fed to tcgen05 as MN-major tiles: a TMA box of 64 rows x 64 channels
// with the 128-byte swizzle IS the canonical MN-major SWIZZLE_128B atom (rows = contraction index, 128-byte rows of 64
// channels, 8-row groups 1024 B apart = SBO; the second 64-channel block LBO = 8192 B further), so the tap shift is the box's
// ROW coordinate -- any integer, negative rows are zero-filled -- and ONE padded copy of the activation serves every tap
// (the K-major version above needs one pre-shifted transposed copy per tap: a box must start 16-byte aligned along the
// contiguous dimension).
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;      // leading byte offset: next 64-element block along M / N
  d |= (uint64_t)(1024 >> 4) << 32;      // stride byte offset: next 8-row group along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor, bf16 x bf16 -> fp32, BOTH operands MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {
  return make_idesc_bf16(M, N) | (1u << 15) | (1u << 16);
}

struct WgShifts {
  int n;
  int shift[kWgMaxTaps];
};

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_mn_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int n_steps,
                     WgShifts taps, float* __restrict__ dW /* [taps.n][128][128], accumulated */) {
  extern __shared__ uint8_t smem_raw[];
  auto& sm = *reinterpret_cast<WgSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s0 = (int)((long long)n_steps * blockIdx.x / gridDim.x);
  const int s1 = (int)((long long)n_steps * (blockIdx.x + 1) / gridDim.x);
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, sm.tmem_base, 0);

  if (warp == 0) {
    if (lane == 0) {
      for (int i = s0; i < s1; ++i) {
        const int n = i - s0, s = n % kWgStages;
        mbar_wait_relaxed(&sm.empty[s], (uint32_t)(((n / kWgStages) & 1) ^ 1));
        mbar_arrive_expect_tx(&sm.full[s], (uint32_t)((1 + taps.n) * kWgTile));
        const int k0 = i * 64;                                   // 64 rows of the contraction per stage
        tma_load_2d(sm.b[s], &tmap_b, 0, k0, &sm.full[s]);
        tma_load_2d(sm.b[s] + kWgTile / 2, &tmap_b, 64, k0, &sm.full[s]);
        for (int t = 0; t < taps.n; ++t) {
          tma_load_2d(sm.a[s][t], &tmap_a, 0, k0 - taps.shift[t], &sm.full[s]);
          tma_load_2d(sm.a[s][t] + kWgTile / 2, &tmap_a, 64, k0 - taps.shift[t], &sm.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16_mn(128, 128);
    for (int i = s0; i < s1; ++i) {
      const int n = i - s0, s = n % kWgStages;
      mbar_wait(&sm.full[s], (uint32_t)((n / kWgStages) & 1));
      tc_fence_after_sync();
      for (int t = 0; t < taps.n; ++t) {
        // descriptor low words (make_desc_mn_sw128) advanced by 16-byte units per K step instead of a 64-bit rebuild per MMA
        constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t a_lo = ((smem_u32(sm.a[s][t]) & 0x3FFFF) >> 4) | ((8192u >> 4) << 16);
        const uint32_t b_lo = ((smem_u32(sm.b[s]) & 0x3FFFF) >> 4) | ((8192u >> 4) << 16);
#pragma unroll
        for (int k = 0; k < 4; ++k)                              // 16 rows of the contraction = two 8-row groups = 2048 B
          if (leader) umma_bf16_lohi(tmem + t * 128, a_lo + k * (2048 >> 4), kHi, b_lo + k * (2048 >> 4), kHi, idesc, (n | k) != 0);
      }
      if (leader) umma_commit(&sm.empty[s]);
    }
    if (leader) umma_commit(&sm.done);
  } else if (s1 > s0) {
    mbar_wait(&sm.done, 0);
    tc_fence_after_sync();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    for (int t = 0; t < taps.n; ++t) {
      float* dst = dW + ((long long)t * 128 + row) * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + t * 128 + c * 32, v);
        tmem_ld_wait(v);
#pragma unroll
        for (int u = 0; u < 32; u += 4)
          atomicAdd(reinterpret_cast<float4*>(dst + c * 32 + u),
                    make_float4(__uint_as_float(v[u]), __uint_as_float(v[u + 1]), __uint_as_float(v[u + 2]), __uint_as_float(v[u + 3])));
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem);
  }
}

// src [R,128] (f32 or bf16) -> dst [Kp][128] bf16, row-major: row(col) as in PadGeom (slot-major, every sequence preceded by
// P zero rows); one warp per destination row
template <bool kBf16>
__global__ void __launch_bounds__(256)
pad_rows_bf16_kernel(const void* __restrict__ src, PadGeom g, __nv_bfloat16* __restrict__ dst) {
  constexpr int kRowsPerWarp = 8;                                  // 8 loads in flight per lane
  const int lane = threadIdx.x & 31;
  const long long col0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * kRowsPerWarp;
  if (col0 >= g.Kp) return;
  long long my_r = -1;                                             // lane i < 8 resolves destination row col0 + i
  {
    const long long col = col0 + lane;
    if (lane < kRowsPerWarp && col < g.base[g.n_slots]) {
      int s = 0;
      while (s + 1 < g.n_slots && g.base[s + 1] <= col) ++s;
      const int W = g.off[s + 1] - g.off[s] + g.P;
      const long long rel = col - g.base[s];
      const int b = (int)(rel / W), t = (int)(rel % W) - g.P;
      if (t >= 0) my_r = (long long)b * g.T + g.off[s] + t;
    }
  }
  uint2 o[kRowsPerWarp];
#pragma unroll
  for (int i = 0; i < kRowsPerWarp; ++i) {
    const long long r = __shfl_sync(0xffffffffu, my_r, i);
    o[i] = make_uint2(0u, 0u);
    if (r >= 0) {
      if (kBf16) {
        o[i] = reinterpret_cast<const uint2*>(src)[r * 32 + lane];
      } else {
        const float4 v = ldg_nc_f4(reinterpret_cast<const float4*>(src) + r * 32 + lane);
        o[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kRowsPerWarp; ++i)
    if (col0 + i < g.Kp) reinterpret_cast<uint2*>(dst)[(col0 + i) * 32 + lane] = o[i];
}

// src [R,128] (f32 or bf16) -> n_shift copies dst_i [128][Kp] bf16 (dst_stride elements apart):
//     dst_i[c][col] = src[row(col - shift_i)][c]   where col - shift_i falls on a position of the SAME sequence, else 0.
// Column layout (slot-major): slot s starts at base[s]; user b's sequence occupies W_s = L_s + P columns, P zero columns
// first (P >= every shift).  A block reads the 64 + max_shift source columns of its tile once.
struct PadShifts {
  int n, max_shift;
  int shift[8];
};
constexpr int kPadHalo = 32;

template <bool kBf16>
__global__ void __launch_bounds__(256)
pad_transpose_bf16_kernel(const void* __restrict__ src, PadGeom g, PadShifts sh, __nv_bfloat16* __restrict__ dst,
                          long long dst_stride) {
  extern __shared__ float pt_smem[];
  float (*tile)[129] = reinterpret_cast<float (*)[129]>(pt_smem);           // [64 + kPadHalo][129]: tile row j = column col0 - kPadHalo + j
  __shared__ long long srow[64 + kPadHalo];
  __shared__ int sstart[64 + kPadHalo];                                     // first column of the sequence window the column lies in
  const int tid = threadIdx.x;
  const long long col0 = (long long)blockIdx.x * 64;
  if (tid < 64 + kPadHalo) {
    const long long col = col0 - kPadHalo + tid;
    long long r = -1;
    int wstart = 0x7fffffff;                                                // no window: every source is refused
    if (col >= 0 && col < g.base[g.n_slots]) {
      int s = 0;
      while (s + 1 < g.n_slots && g.base[s + 1] <= col) ++s;
      const int W = g.off[s + 1] - g.off[s] + g.P;
      const long long rel = col - g.base[s];
      const int b = (int)(rel / W), t = (int)(rel % W) - g.P;
      if (t >= 0) r = (long long)b * g.T + g.off[s] + t;
      wstart = (int)(col - rel % W - (col0 - kPadHalo));                   // tile-row index of the window's first column
    }
    srow[tid] = r;
    sstart[tid] = wstart;
  }
  __syncthreads();
  const int j_lo = kPadHalo - sh.max_shift;
  for (int i = tid; i < (64 + kPadHalo) * 32; i += 256) {
    const int j = i >> 5, q = i & 31;
    if (j < j_lo) continue;
    const long long r = srow[j];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 0) {
      if (kBf16) {
        const uint2 u = reinterpret_cast<const uint2*>(src)[r * 32 + q];
        v = make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
      } else {
        v = reinterpret_cast<const float4*>(src)[r * 32 + q];
      }
    }
    tile[j][q * 4 + 0] = v.x; tile[j][q * 4 + 1] = v.y; tile[j][q * 4 + 2] = v.z; tile[j][q * 4 + 3] = v.w;
  }
  __syncthreads();
  for (int si = 0; si < sh.n; ++si) {
    const int s_ = sh.shift[si];
    for (int i = tid; i < 128 * 8; i += 256) {
      const int ch = i >> 3, seg = i & 7;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int jd = kPadHalo + seg * 8 + e;            // destination column, as a tile row
        const int js = jd - s_;                           // source column
        // same sequence window: the source must not lie before the window of the destination column
        v[e] = js >= sstart[jd] ? tile[js][ch] : 0.f;      // (the window may start left of the tile: negative)
      }
      *reinterpret_cast<uint4*>(dst + si * dst_stride + (long long)ch * g.Kp + col0 + seg * 8) =
          make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
  }
}

}  // namespace

PadGeom make_pad_geom(const SlotTable& slots, int B, int T, int P) {
  PadGeom g{};
  g.n_slots = slots.n; g.B = B; g.T = T; g.P = P;
  long long base = 0;
  for (int s = 0; s <= slots.n; ++s) {
    g.off[s] = slots.off[s];
    g.base[s] = base;
    if (s < slots.n) base += (long long)B * (slots.off[s + 1] - slots.off[s] + P);
  }
  g.Kp = (base + 63) / 64 * 64;
  return g;
}

int32_t pad_transpose_bf16(const void* src, bool src_bf16, const PadGeom& g, const int* shifts, int n_shifts, void* dst,
                           long long dst_stride, cudaStream_t st) {
  PadShifts sh{};
  sh.n = n_shifts;
  for (int i = 0; i < n_shifts; ++i) {
    sh.shift[i] = shifts[i];
    if (shifts[i] > sh.max_shift) sh.max_shift = shifts[i];
  }
  if (n_shifts < 1 || n_shifts > 8 || sh.max_shift > kPadHalo || sh.max_shift > g.P) {
    set_error("pad_transpose: %d shifts, largest %d (halo %d, pad %d)", n_shifts, sh.max_shift, kPadHalo, g.P);
    return HTCN_ERR_INVALID;
  }
  const int grid = (int)(g.Kp / 64);
  const size_t smem = sizeof(float) * (64 + kPadHalo) * 129;
  if (src_bf16) {
    HTCN_CUDA(cudaFuncSetAttribute(pad_transpose_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pad_transpose_bf16_kernel<true><<<grid, 256, smem, st>>>(src, g, sh, reinterpret_cast<__nv_bfloat16*>(dst), dst_stride);
  } else {
    HTCN_CUDA(cudaFuncSetAttribute(pad_transpose_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pad_transpose_bf16_kernel<false><<<grid, 256, smem, st>>>(src, g, sh, reinterpret_cast<__nv_bfloat16*>(dst), dst_stride);
  }
  HTCN_LAUNCH_CHECK("pad_transpose_bf16_kernel");
  return HTCN_OK;
}

// dW[t][cin][cout] += sum_col aT_t[cin][col] * bT[cout][col]   (aT_t = aT + t * a_stride: tap t's pre-shifted copy; bf16
// [128][Kp] from pad_transpose_bf16)
int32_t wgrad_bf16(const void* aT, long long a_stride, const void* bT, long long Kp, int n_taps, float* dW, cudaStream_t st) {
  CUtensorMap tb;
  int32_t rc = make_tmap_bf16(&tb, bT, 128, (uint32_t)Kp, (uint32_t)Kp, 64, 128, 128);
  if (rc) return rc;
  const size_t smem = sizeof(WgSmem) + 1024;
  HTCN_CUDA(cudaFuncSetAttribute(wgrad_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_steps = (int)(Kp / 64);
  const int grid = n_steps < 148 ? n_steps : 148;
  for (int t0 = 0; t0 < n_taps; t0 += kWgMaxTaps) {
    WgTaps taps{};
    taps.n = n_taps - t0 < kWgMaxTaps ? n_taps - t0 : kWgMaxTaps;
    CUtensorMap ta[kWgMaxTaps];
    for (int t = 0; t < kWgMaxTaps; ++t) {
      const int tt = t < taps.n ? t0 + t : t0;
      rc = make_tmap_bf16(&ta[t], reinterpret_cast<const __nv_bfloat16*>(aT) + tt * a_stride, 128, (uint32_t)Kp, (uint32_t)Kp, 64, 128, 128);
      if (rc) return rc;
    }
    wgrad_bf16_kernel<<<grid, kWgThreads, smem, st>>>(ta[0], ta[1], ta[2], tb, n_steps, taps, dW + (long long)t0 * 128 * 128);
    HTCN_LAUNCH_CHECK("wgrad_bf16_kernel");
  }
  return HTCN_OK;
}


int32_t pad_rows_bf16(const void* src, bool src_bf16, const PadGeom& g, void* dst, cudaStream_t st) {
  const int grid = (int)((g.Kp + 63) / 64);
  if (src_bf16) pad_rows_bf16_kernel<true><<<grid, 256, 0, st>>>(src, g, reinterpret_cast<__nv_bfloat16*>(dst));
  else pad_rows_bf16_kernel<false><<<grid, 256, 0, st>>>(src, g, reinterpret_cast<__nv_bfloat16*>(dst));
  HTCN_LAUNCH_CHECK("pad_rows_bf16_kernel");
  return HTCN_OK;
}

// dW[t][cin][cout] += sum_r aP[r - shifts[t]][cin] * bP[r][cout]   (aP, bP: bf16 [Kp][128] from pad_rows_bf16)
int32_t wgrad_mn_bf16(const void* aP, const void* bP, long long Kp, const int* shifts, int n_taps, float* dW, cudaStream_t st) {
  CUtensorMap ta, tb;
  int32_t rc = make_tmap_bf16(&ta, aP, (uint64_t)Kp, 128, 128, 64, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tb, bP, (uint64_t)Kp, 128, 128, 64, 64, 128);
  if (rc) return rc;
  const size_t smem = sizeof(WgSmem) + 1024;
  HTCN_CUDA(cudaFuncSetAttribute(wgrad_mn_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_steps = (int)(Kp / 64);
  const int grid = n_steps < 148 ? n_steps : 148;
  for (int t0 = 0; t0 < n_taps; t0 += kWgMaxTaps) {
    WgShifts taps{};
    taps.n = n_taps - t0 < kWgMaxTaps ? n_taps - t0 : kWgMaxTaps;
    for (int t = 0; t < taps.n; ++t) taps.shift[t] = shifts[t0 + t];
    wgrad_mn_bf16_kernel<<<grid, kWgThreads, smem, st>>>(ta, tb, n_steps, taps, dW + (long long)t0 * 128 * 128);
    HTCN_LAUNCH_CHECK("wgrad_mn_bf16_kernel");
  }
  return HTCN_OK;
}

}  // namespace htcn

// test hooks (not part of include/htcn.h): the two stages of the tensor-core weight gradient on their own
extern "C" int32_t htcn_debug_pad_transpose(const void* src, int32_t src_bf16, const int32_t* slot_off_host, int32_t B, int32_t T,
                                            int32_t S, int32_t P, const int32_t* shifts_host, int32_t n_shifts, void* dst,
                                            int64_t* kp_out, void* stream) {
  using namespace htcn;
  SlotTable slots;
  slots.n = S;
  for (int i = 0; i <= S; ++i) slots.off[i] = slot_off_host[i];
  const PadGeom g = make_pad_geom(slots, B, T, P);
  if (kp_out) *kp_out = g.Kp;
  if (!dst) return HTCN_OK;
  return pad_transpose_bf16(src, src_bf16 != 0, g, shifts_host, n_shifts, dst, 128 * g.Kp, as_stream(stream));
}
extern "C" int32_t htcn_debug_wgrad(const void* aT, const void* bT, int64_t Kp, int32_t n_taps, float* dW, void* stream) {
  return htcn::wgrad_bf16(aT, 128 * Kp, bT, Kp, n_taps, dW, htcn::as_stream(stream));
}

extern "C" int32_t htcn_debug_pad_rows(const void* src, int32_t src_bf16, const int32_t* slot_off_host, int32_t B, int32_t T,
                                       int32_t S, int32_t P, void* dst, void* stream) {
  using namespace htcn;
  SlotTable slots;
  slots.n = S;
  for (int i = 0; i <= S; ++i) slots.off[i] = slot_off_host[i];
  const PadGeom g = make_pad_geom(slots, B, T, P);
  return pad_rows_bf16(src, src_bf16 != 0, g, dst, as_stream(stream));
}
extern "C" int32_t htcn_debug_wgrad_mn(const void* aP, const void* bP, int64_t Kp, const int32_t* shifts_host, int32_t n_taps,
                                       float* dW, void* stream) {
  return htcn::wgrad_mn_bf16(aP, bP, Kp, shifts_host, n_taps, dW, htcn::as_stream(stream));
}
