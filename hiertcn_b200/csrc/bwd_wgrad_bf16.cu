// Weight gradients of the conv stack on the 5th-gen tensor cores (bf16 tier of the training step).
//
//     dW_l[tap][cin][cout] += sum_r h_l[src(r, tap)][cin] * dp[r][cout],   src(r, tap) = r - (K-1-tap) d, zero before the
//                                                                          start of r's sequence (customized_tcn_cell.py:46-49)
// is a "TN" GEMM whose contraction runs over the B*T positions (819 200 at config 5) with M = N = 128: 2.1 GFLOP per tap
// that the fp32 split-K kernel (sgemm_tn_atomic) runs at FFMA speed -- 8.3 ms of the 49 ms step.  Here both operands are
// first written TRANSPOSED and ZERO-PADDED as bf16 [128][Kp] (pad_transpose_bf16: every sequence is preceded by P = the
// largest shift zero columns, so a shifted read can never reach the previous sequence -- the same physical-padding idea as
// the forward conv kernel), which makes every tap a plain K-major x K-major tcgen05 GEMM whose A tile is the TMA box at
// column k0 - shift (negative coordinates are zero-filled).  One CTA accumulates up to 3 taps (3 x 128 TMEM columns) over
// its share of the columns and adds its partial to dW with float4 atomics.
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
  int shift[kWgMaxTaps];
};

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int n_steps,
                  WgTaps taps, float* __restrict__ dW /* [taps.n][128][128], accumulated */) {
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
        const int k0 = i * 64;
        tma_load_2d(sm.b[s], &tmap_b, k0, 0, &sm.full[s]);
        for (int t = 0; t < taps.n; ++t) tma_load_2d(sm.a[s][t], &tmap_a, k0 - taps.shift[t], 0, &sm.full[s]);
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

// src [R,128] (f32 or bf16) -> dst [128][Kp] bf16: dst[c][col(r)] = src[r][c], zero in the pad columns.  Column layout
// (slot-major): slot s starts at base[s]; user b's sequence occupies W_s = L_s + P columns, P zero columns first.
template <bool kBf16>
__global__ void __launch_bounds__(256)
pad_transpose_bf16_kernel(const void* __restrict__ src, PadGeom g, __nv_bfloat16* __restrict__ dst) {
  __shared__ float tile[64][129];
  __shared__ long long srow[64];
  const int tid = threadIdx.x;
  const long long col0 = (long long)blockIdx.x * 64;
  if (tid < 64) {
    const long long col = col0 + tid;
    long long r = -1;
    if (col < g.base[g.n_slots]) {
      int s = 0;
      while (s + 1 < g.n_slots && g.base[s + 1] <= col) ++s;
      const int W = g.off[s + 1] - g.off[s] + g.P;
      const long long rel = col - g.base[s];
      const int b = (int)(rel / W), t = (int)(rel % W) - g.P;
      if (t >= 0) r = (long long)b * g.T + g.off[s] + t;
    }
    srow[tid] = r;
  }
  __syncthreads();
  for (int i = tid; i < 64 * 32; i += 256) {
    const int c = i >> 5, q = i & 31;
    const long long r = srow[c];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 0) {
      if (kBf16) {
        const uint2 u = reinterpret_cast<const uint2*>(src)[r * 32 + q];
        v = make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
      } else {
        v = reinterpret_cast<const float4*>(src)[r * 32 + q];
      }
    }
    tile[c][q * 4 + 0] = v.x; tile[c][q * 4 + 1] = v.y; tile[c][q * 4 + 2] = v.z; tile[c][q * 4 + 3] = v.w;
  }
  __syncthreads();
  for (int i = tid; i < 128 * 8; i += 256) {
    const int ch = i >> 3, seg = i & 7;
    const uint4 o = make_uint4(pack_bf16x2(tile[seg * 8 + 0][ch], tile[seg * 8 + 1][ch]), pack_bf16x2(tile[seg * 8 + 2][ch], tile[seg * 8 + 3][ch]),
                               pack_bf16x2(tile[seg * 8 + 4][ch], tile[seg * 8 + 5][ch]), pack_bf16x2(tile[seg * 8 + 6][ch], tile[seg * 8 + 7][ch]));
    *reinterpret_cast<uint4*>(dst + (long long)ch * g.Kp + col0 + seg * 8) = o;
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

int32_t pad_transpose_bf16(const void* src, bool src_bf16, const PadGeom& g, void* dst, cudaStream_t st) {
  const int grid = (int)(g.Kp / 64);
  if (src_bf16) pad_transpose_bf16_kernel<true><<<grid, 256, 0, st>>>(src, g, reinterpret_cast<__nv_bfloat16*>(dst));
  else pad_transpose_bf16_kernel<false><<<grid, 256, 0, st>>>(src, g, reinterpret_cast<__nv_bfloat16*>(dst));
  HTCN_LAUNCH_CHECK("pad_transpose_bf16_kernel");
  return HTCN_OK;
}

// dW[t][cin][cout] += sum_col aT[cin][col - shift[t]] * bT[cout][col]     (aT, bT: bf16 [128][Kp] from pad_transpose_bf16)
int32_t wgrad_bf16(const void* aT, const void* bT, long long Kp, const int* shifts, int n_taps, float* dW, cudaStream_t st) {
  CUtensorMap ta, tb;
  int32_t rc = make_tmap_bf16(&ta, aT, 128, (uint32_t)Kp, (uint32_t)Kp, 64, 128, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tb, bT, 128, (uint32_t)Kp, (uint32_t)Kp, 64, 128, 128);
  if (rc) return rc;
  const size_t smem = sizeof(WgSmem) + 1024;
  HTCN_CUDA(cudaFuncSetAttribute(wgrad_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_steps = (int)(Kp / 64);
  const int grid = n_steps < 148 ? n_steps : 148;
  for (int t0 = 0; t0 < n_taps; t0 += kWgMaxTaps) {
    WgTaps taps{};
    taps.n = n_taps - t0 < kWgMaxTaps ? n_taps - t0 : kWgMaxTaps;
    for (int t = 0; t < taps.n; ++t) taps.shift[t] = shifts[t0 + t];
    wgrad_bf16_kernel<<<grid, kWgThreads, smem, st>>>(ta, tb, n_steps, taps, dW + (long long)t0 * 128 * 128);
    HTCN_LAUNCH_CHECK("wgrad_bf16_kernel");
  }
  return HTCN_OK;
}

}  // namespace htcn
