// K3 (bf16 tier), wavefront variant of the users-on-N kernel (k3_gru_t.cu): layer 0 of step t+1 runs beside layer 1 of step t.
//
// In k3_gru_t.cu a step is four dependent phases (gates 0 -> candidate 0 -> gates 1 -> candidate 1); each costs the phase's MMAs
// (64 clk per M = 128, N = 32, K = 16 tcgen05.mma: the A operand is read from shared memory) PLUS an epilogue round trip of
// ~2000 clk (commit -> mbarrier -> tcgen05.ld -> sigmoid / tanh -> operand stores -> fence -> mbarrier -> issue), so the tensor
// pipe idles half of the time (ncu: 41 % active).  But layer 0 of step t+1 needs only h0(t) and x(t+1) -- not layer 1 of step t.
// Two independent chains exist at any time; this kernel gives each its own 8 epilogue warps and interleaves their products in one
// MMA issue stream, so an epilogue round trip of one chain hides behind the products of the other:
//
//   iteration t = 0..S          (layer 0 works on step t, layer 1 on step t-1)
//     1. wait L0c(t)   : G0r(t) G0u(t)                       [X | H0]      -> acc r0, u0
//     2. wait L1c(t-1) : G1r(t-1) G1u(t-1)                   [U0 | H1]     -> acc r1, u1
//                        SBa(t)  += H0 W_in[D:D+128]          (sbias[t],   first half of K)
//                        SBb(t-1) += H1 W_in[D+128:]          (sbias[t-1], second half; drained by the layer-1 warps)
//     3. wait R0(t)    : C0(t)                               [X | R0]      -> acc c0
//     4. wait T1(t-1)  : C1(t-1)                             [U0 | T1]     -> acc c1
//   layer-0 warps:  R0 <- r*h0 (publishes R0) ... h0' ; U0[t&1] <- h0' ; H0 <- m*h0' ; X <- x(t+1) (publishes L0c)
//   layer-1 warps:  T1 <- r*h1 (publishes T1) ... sbias[t] out ... H1 <- m*h1' (publishes L1c)
// U0 (the unmasked h0' layer 1 reads) is double-buffered: C1(t-1) is still in flight when layer 0 writes h0'(t).  sbias[t] =
// [h0(t-1) | h1(t-1)] W_in[D:] is split along K because its two halves exist in different iterations; two accumulators alternate.
// Same arithmetic as the other bf16 GRU kernels (bit-equal results: tests/test_gpu_kernels.py::test_k3_cluster_variants_agree).
// Reference: customed_gru_cell.py:309-337 per layer, :1050-1073 stacking; model_hier.py:54-55,91,93.
#include <cstdlib>

#include "common.cuh"
#include "sm100.cuh"

namespace htcn {
using namespace sm100;

int32_t k3_prepare_rep(const float* const* gate_w, const float* const* gate_b, const float* const* cand_w,
                       const float* const* cand_b, const float* w_in_state, float* scratch, cudaStream_t st,
                       const uint8_t** w_out, float** bias_out);

namespace k3w {
constexpr int kNU = 32;                           // users per CTA = MMA N
constexpr int kSub = 128 * 64 * 2;                // weight sub-tile: 128 hidden units x 64 k, bf16, 128-byte swizzle = 16 KB
constexpr int kUPT = 16;                          // users per epilogue thread
constexpr int kGroupWarps = 8;                    // per layer: 4 TMEM lane quarters x 2 user groups of 16
constexpr int kEpiWarps = 2 * kGroupWarps;
constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;
constexpr int kThreads = 32 * (kEpiWarps + 2);
constexpr int kChunkStride = kNU * 16 + 16;       // see k3_gru_t.cu: 16 bytes of padding keep a warp's 2-byte stores conflict-free
constexpr int kSlotBytes = 16 * kChunkStride;
enum Slot { kX = 0, kH0, kR0, kU0a, kU0b, kH1, kT1, kSlots };
// TMEM columns (32 each): r0 u0 c0 | r1 u1 c1 | sbias accumulators of even / odd steps
constexpr uint32_t kColR = 0, kColU = 32, kColC = 64, kColLayer = 96, kColSb = 192, kTmemCols = 256;
// weight tiles ([128 n][128 k], k3_prepare_weights order): sbias k0 k1 | layer 0: r k0 k1, u k0 k1, cand k0 k1 | layer 1: ...
constexpr int kTileSb = 0, kTileG0r = 2, kTileG0u = 4, kTileC0 = 6, kTileG1r = 8, kTileG1u = 10, kTileC1 = 12;

template <int kRes, int kStages>
struct alignas(1024) Smem {
  uint8_t res[kRes > 0 ? kRes : 1][kSub];         // sub-tiles of the layer-0 gate kernel, resident for the whole call
  uint8_t ring[kStages][kSub];
  alignas(16) uint8_t act[kSlots][kSlotBytes];
  uint64_t w_full[kStages], w_empty[kStages], res_full;
  uint64_t acc_r[2], acc_u[2], acc_c[2], acc_sb;  // MMA warp -> epilogue warps of layer l
  uint64_t pub_g[2], pub_c[2];                    // epilogue warps of layer l -> MMA warp: r*h stored / step finished
  uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t smem_addr, uint32_t lbo) {   // 8-row groups 128 B apart
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }
__device__ __forceinline__ void sts_bf16(uint32_t saddr, float v) {
  const unsigned short b = __bfloat16_as_ushort(__float2bfloat16_rn(v));
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(saddr), "h"(b) : "memory");
}
__host__ __device__ constexpr bool resident(int tile, int hf, int n_res) {
  return tile >= kTileG0r && tile < kTileC0 && (tile - kTileG0r) * 2 + hf < n_res;
}
}  // namespace k3w

template <int kRes, int kStages>
__global__ void __launch_bounds__(k3w::kThreads, 1)
k3_gru_bf16_w(const uint8_t* __restrict__ w_img, const float* __restrict__ yp, const float* __restrict__ mask,
              const float* __restrict__ state_in, const float* __restrict__ bias_all /* [bg0 256][bc0 128][bg1 256][bc1 128] */,
              int B, int S, int do_sbias, float* __restrict__ state_pre, float* __restrict__ sbias,
              float* __restrict__ state_out) {
  using namespace k3w;
  extern __shared__ uint8_t smem_raw[];
  auto& sm = *reinterpret_cast<Smem<kRes, kStages>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.w_full[s], 1);
      mbar_init(&sm.w_empty[s], 1);
    }
    mbar_init(&sm.res_full, 1);
    for (int l = 0; l < 2; ++l) {
      mbar_init(&sm.acc_r[l], 1);
      mbar_init(&sm.acc_u[l], 1);
      mbar_init(&sm.acc_c[l], 1);
      mbar_init(&sm.pub_g[l], kGroupWarps);      // one arrive per epilogue warp of the layer
      mbar_init(&sm.pub_c[l], kGroupWarps);
    }
    mbar_init(&sm.acc_sb, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<kTmemCols>(&sm.tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, sm.tmem_base, 0);

  if (warp == kProducerWarp) {
    // ===================== weight producer: same tile order as the MMA issuer =====================
    if (lane == 0) {
      if (kRes > 0) {
        mbar_arrive_expect_tx(&sm.res_full, kRes * kSub);
        for (int i = 0; i < kRes; ++i) bulk_load_1d(sm.res[i], w_img + (size_t)((kTileG0r + (i >> 1)) * 2 + (i & 1)) * kSub, kSub, &sm.res_full);
      }
      long long n = 0;
      auto emit = [&](int tile) {
        for (int hf = 0; hf < 2; ++hf) {
          if (resident(tile, hf, kRes)) continue;
          const int st = (int)(n % kStages);
          mbar_wait_relaxed(&sm.w_empty[st], (uint32_t)(((n / kStages) & 1) ^ 1));
          mbar_arrive_expect_tx(&sm.w_full[st], kSub);
          bulk_load_1d(sm.ring[st], w_img + (size_t)(tile * 2 + hf) * kSub, kSub, &sm.w_full[st]);
          ++n;
        }
      };
      for (int t = 0; t <= S; ++t) {
        if (t < S) { emit(kTileG0r); emit(kTileG0r + 1); emit(kTileG0u); emit(kTileG0u + 1); }
        if (t >= 1) { emit(kTileG1r); emit(kTileG1r + 1); emit(kTileG1u); emit(kTileG1u + 1); }
        if (do_sbias && t < S) emit(kTileSb);
        if (do_sbias && t >= 1) emit(kTileSb + 1);
        if (t < S) { emit(kTileC0); emit(kTileC0 + 1); }
        if (t >= 1) { emit(kTileC1); emit(kTileC1 + 1); }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(128, kNU);
    const uint32_t act0 = smem_u32(sm.act);
    long long n = 0;
    // one weight tile = K 128 = one operand slot = two sub-tiles of 4 MMAs
    auto tile_mma = [&](int tile, int slot, uint32_t d_col, bool fresh) {
      const uint32_t slot_addr = act0 + (uint32_t)slot * kSlotBytes;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const bool res = resident(tile, hf, kRes);
        uint32_t w_base;
        int st = 0;
        if (res) {
          w_base = smem_u32(sm.res[res ? (tile - kTileG0r) * 2 + hf : 0]);
        } else {
          st = (int)(n % kStages);
          mbar_wait(&sm.w_full[st], (uint32_t)((n / kStages) & 1));
          tc_fence_after_sync();
          w_base = smem_u32(sm.ring[st]);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t da = make_desc_k_sw128(w_base + kk * 32);
          const uint64_t db = make_desc_nosw(slot_addr + (uint32_t)(hf * 8 + kk * 2) * kChunkStride, kChunkStride);
          if (leader) umma_bf16(tmem + d_col, da, db, idesc, !(fresh && hf == 0 && kk == 0));
        }
        if (!res) {
          if (leader) umma_commit(&sm.w_empty[st]);
          ++n;
        }
      }
    };
    auto wait_pub = [&](uint64_t* bar, int phase) {
      mbar_wait(bar, (uint32_t)(phase & 1));
      tc_fence_after_sync();
    };
    if (kRes > 0) {
      mbar_wait(&sm.res_full, 0);
      tc_fence_after_sync();
    }
    for (int t = 0; t <= S; ++t) {
      const int u0_prev = kU0a + ((t - 1) & 1);                       // the U0 buffer of step t-1
      wait_pub(&sm.pub_c[0], t);                                      // 1. X(t), H0(t-1), U0(t-1)
      if (t < S) {
        tile_mma(kTileG0r, kX, kColR, true);
        tile_mma(kTileG0r + 1, kH0, kColR, false);
        if (leader) umma_commit(&sm.acc_r[0]);
        tile_mma(kTileG0u, kX, kColU, true);
        tile_mma(kTileG0u + 1, kH0, kColU, false);
        if (leader) umma_commit(&sm.acc_u[0]);
      }
      if (t >= 1) {
        wait_pub(&sm.pub_c[1], t - 1);                                // 2. H1(t-2); the layer-1 warps have drained sbias[t-2]
        tile_mma(kTileG1r, u0_prev, kColLayer + kColR, true);
        tile_mma(kTileG1r + 1, kH1, kColLayer + kColR, false);
        if (leader) umma_commit(&sm.acc_r[1]);
        tile_mma(kTileG1u, u0_prev, kColLayer + kColU, true);
        tile_mma(kTileG1u + 1, kH1, kColLayer + kColU, false);
        if (leader) umma_commit(&sm.acc_u[1]);
      }
      if (do_sbias) {
        if (t < S) tile_mma(kTileSb, kH0, kColSb + 32 * (t & 1), true);
        if (t >= 1) {
          tile_mma(kTileSb + 1, kH1, kColSb + 32 * ((t - 1) & 1), false);
          if (leader) umma_commit(&sm.acc_sb);
        }
      }
      if (t < S) {
        wait_pub(&sm.pub_g[0], t);                                    // 3. R0 = r * h0
        tile_mma(kTileC0, kX, kColC, true);
        tile_mma(kTileC0 + 1, kR0, kColC, false);
        if (leader) umma_commit(&sm.acc_c[0]);
      }
      if (t >= 1) {
        wait_pub(&sm.pub_g[1], t - 1);                                // 4. T1 = r * h1
        tile_mma(kTileC1, u0_prev, kColLayer + kColC, true);
        tile_mma(kTileC1 + 1, kT1, kColLayer + kColC, false);
        if (leader) umma_commit(&sm.acc_c[1]);
      }
    }
  } else {
    // ===================== epilogue: thread = hidden unit j of layer l for 16 users =====================
    const int l = warp >> 3, quarter = warp & 3, ug = (warp >> 2) & 1;
    const int j = quarter * 32 + lane;                            // hidden unit = TMEM lane
    const int u0 = ug * kUPT;                                     // first user (row of the operand slots) of this thread
    const long long b0 = (long long)blockIdx.x * kNU + u0;
    const int n_ok = (int)max(0LL, min((long long)kUPT, (long long)B - b0));
    const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)u0 + (uint32_t)l * kColLayer;
    const uint32_t t_sb = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)u0 + kColSb;
    // shared address of (user u0, k = j) in slot 0; users are 16 B apart, slots kSlotBytes apart
    const uint32_t a0 = smem_u32(sm.act) + (uint32_t)(j >> 3) * kChunkStride + (uint32_t)u0 * 16 + (uint32_t)(j & 7) * 2;
    const float bgr = __ldg(bias_all + l * 384 + j), bgu = __ldg(bias_all + l * 384 + 128 + j),
                bcc = __ldg(bias_all + l * 384 + 256 + j);
    float h[kUPT], m[kUPT], u[kUPT];
    uint32_t xn[kUPT / 2];                                        // layer 0: the next step's input, already bf16
    auto put = [&](int slot, int i, float v) { sts_bf16(a0 + (uint32_t)slot * kSlotBytes + (uint32_t)i * 16, v); };
    auto put_raw = [&](int slot, int i, uint32_t b16) {
      asm volatile("st.shared.b16 [%0], %1;" ::"r"(a0 + (uint32_t)slot * kSlotBytes + (uint32_t)i * 16), "h"((unsigned short)b16) : "memory");
    };
    // every lane orders its own operand stores before the async proxy, then one lane per warp arrives
    auto publish = [&](uint64_t* bar) {
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    auto wait_acc = [&](uint64_t* bar, int phase) {
      mbar_wait(bar, (uint32_t)(phase & 1));
      tc_fence_after_sync();
    };
    auto ld_acc = [&](uint32_t taddr, float (&v)[kUPT]) {
      uint32_t r[16];
      tmem_ld_32x16(taddr, r);
      tmem_ld_wait(r);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
    };
    auto load_x = [&](int s) {                                    // x(s) of this thread's users -> xn (zeros past S / past B)
#pragma unroll
      for (int i = 0; i < kUPT; i += 2) {
        const float a = (i < n_ok && s < S) ? __ldg(yp + ((long long)s * B + b0 + i) * kDim + j) : 0.f;
        const float b = (i + 1 < n_ok && s < S) ? __ldg(yp + ((long long)s * B + b0 + i + 1) * kDim + j) : 0.f;
        xn[i / 2] = pack_bf16x2(a, b);
      }
    };
    // prologue: state (and the first input) -> registers and operand slots
#pragma unroll
    for (int i = 0; i < kUPT; ++i) {
      h[i] = i < n_ok ? __ldg(state_in + (b0 + i) * 256 + l * 128 + j) : 0.f;
      put(l == 0 ? kH0 : kH1, i, h[i]);
    }
    if (l == 0) {
      load_x(0);
#pragma unroll
      for (int i = 0; i < kUPT; ++i) put_raw(kX, i, (i & 1) ? (xn[i / 2] >> 16) : (xn[i / 2] & 0xffffu));
    }
    publish(&sm.pub_c[l]);
    for (int s = 0; s < S; ++s) {
      float v[kUPT];
      wait_acc(&sm.acc_r[l], s);                                    // ---- E_g: R0 / T1 <- r * h
      ld_acc(t_lane + kColR, v);
#pragma unroll
      for (int i = 0; i < kUPT; ++i) put(l == 0 ? kR0 : kT1, i, sigmoid_fast(v[i] + bgr) * h[i]);
      publish(&sm.pub_g[l]);
      // ---- under the candidate product: emit the state before the step, fetch the mask and (layer 0) the next input.  NOT at
      // the top of the step: a fence.proxy.async (every publish) waits for the thread's outstanding global loads (~1500 clk)
      if (state_pre) {
#pragma unroll
        for (int i = 0; i < kUPT; ++i)
          if (i < n_ok) state_pre[((long long)s * B + b0 + i) * 256 + l * 128 + j] = h[i];
      }
#pragma unroll
      for (int i = 0; i < kUPT; ++i) m[i] = i < n_ok ? __ldg(mask + (long long)s * B + b0 + i) : 0.f;
      if (l == 0) load_x(s + 1);
      wait_acc(&sm.acc_u[l], s);                                    // under the candidate product: the update gate
      ld_acc(t_lane + kColU, v);
#pragma unroll
      for (int i = 0; i < kUPT; ++i) u[i] = sigmoid_fast(v[i] + bgu);
      if (l == 1 && do_sbias) {                                     // sbias[s] is complete one iteration after its first half
        wait_acc(&sm.acc_sb, s);
        ld_acc(t_sb + 32 * (s & 1), v);
#pragma unroll
        for (int i = 0; i < kUPT; ++i)
          if (i < n_ok) sbias[((long long)s * B + b0 + i) * kDim + j] = v[i];
      }
      wait_acc(&sm.acc_c[l], s);                                    // ---- E_c: h' = u*h + (1-u)*c
      ld_acc(t_lane + kColC, v);
#pragma unroll
      for (int i = 0; i < kUPT; ++i) {
        const float c = tanh_fast(v[i] + bcc);
        const float o = fmaf(u[i], h[i] - c, c);                    // UNMASKED: the input of the layer above
        h[i] = m[i] * o;                                            // state *= mask (model_hier.py:93)
        if (l == 0) {
          put(kU0a + (s & 1), i, o);
          put(kH0, i, h[i]);
          put_raw(kX, i, (i & 1) ? (xn[i / 2] >> 16) : (xn[i / 2] & 0xffffu));
        } else {
          put(kH1, i, h[i]);
        }
      }
      publish(&sm.pub_c[l]);
    }
#pragma unroll
    for (int i = 0; i < kUPT; ++i)
      if (i < n_ok) state_out[(b0 + i) * 256 + l * 128 + j] = h[i];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem);
  }
}

template <int kRes, int kStages>
static int32_t launch_w(const uint8_t* tw, const float* yp, const float* mask, const float* state_in, const float* bias_dev,
                        int B, int S, float* state_pre, float* sbias, float* state_out, cudaStream_t st) {
  using Sm = k3w::Smem<kRes, kStages>;
  const size_t smem = sizeof(Sm) + 1024;
  static_assert(sizeof(Sm) + 1024 <= 232448, "shared-memory budget of one CTA");
  auto kern = k3_gru_bf16_w<kRes, kStages>;
  HTCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<ceil_div(B, k3w::kNU), k3w::kThreads, smem, st>>>(tw, yp, mask, state_in, bias_dev, B, S, sbias != nullptr,
                                                         state_pre, sbias, state_out);
  HTCN_LAUNCH_CHECK("k3_gru_bf16_w");
  return HTCN_OK;
}

// variant 0: the r half of the layer-0 gate kernel resident (64 KB), 6-stage ring; 1: nothing resident, 9-stage ring
int32_t gru_sessions_bf16_w(const float* yp, const float* mask, const float* state_in, const float* const* gate_w,
                            const float* const* gate_b, const float* const* cand_w, const float* const* cand_b,
                            const float* w_in_state, int B, int S, float* state_pre, float* sbias, float* state_out,
                            float* scratch, int variant, cudaStream_t st) {
  float* bias_dev;
  const uint8_t* tw;
  int32_t rc = k3_prepare_rep(gate_w, gate_b, cand_w, cand_b, w_in_state, scratch, st, &tw, &bias_dev);
  if (rc) return rc;
  if (variant == 1) return launch_w<0, 9>(tw, yp, mask, state_in, bias_dev, B, S, state_pre, sbias, state_out, st);
  return launch_w<4, 6>(tw, yp, mask, state_in, bias_dev, B, S, state_pre, sbias, state_out, st);
}

}  // namespace htcn
