// K3 (bf16 tier), users-on-N variant: GRU over sessions with NO exchange between CTAs.
//
// The other two tensor-core GRU kernels put the users on the MMA M axis (M = 128 users per CTA or per 4-CTA cluster), so a
// batch of 4096 users either runs on 32 SMs (k3_gru_bf16.cu) or has to split the hidden columns over a cluster and exchange
// every h' slice through distributed shared memory (k3_gru_cluster.cu: 40 dependent phases of ~3.4 us, most of it the
// exchange).  Here the roles of the operands are swapped:
//   D[hidden unit, user] = W^T[hidden unit, k] * act[user, k]^T
// A = a 128-row weight tile (K-major, 128-byte swizzle, the [n][k] tiles of k3_prepare_weights), B = the activations of
// the CTA's own kNU = 32 users (N = 32; no-swizzle K-major, 16-byte K chunks), the accumulator is TMEM lane = hidden unit,
// column = user.  A CTA owns ALL 128 hidden units of its 32 users, so 128 CTAs work on 4096 users and no h' ever leaves the
// SM: the fp32 state of (unit j, 16 users) lives in the registers of one epilogue thread, which writes the bf16 copies the
// next product reads straight into its own CTA's operand slots.
//
// The price is the weights: all five matrices of a step (448 KB bf16) cannot be resident.  The layer-0 gate kernel (128 KB)
// stays in shared memory for the whole call; the other 20 sub-tiles of 16 KB stream from L2 through a TMA ring every step
// -- they do not depend on the recurrence, so the producer warp runs ahead of it.
//
// Operand slots (bf16 [32 users][128 k] each): X | H0 | T0 | H1 | T1.   Step s (customed_gru_cell.py:309-337 per layer,
// :1050-1073 stacking; model_hier.py:54-55,91,93):
//   G0 : [r|u] = sigmoid([X | H0] Wg0 + bg0)           E: T0 <- r * h0                       (the u half runs under this epilogue)
//   C0 : c = tanh([X | T0] Wc0 + bc0)                  E: h0' = u*h0 + (1-u)*c ; T0 <- h0'
//   SB : sbias[s] = [H0 | H1] @ W_in[D:]               runs under E_c0; then H0 <- m*h0', X <- x_{s+1}
//   G1 : [r|u] = sigmoid([T0 | H1] Wg1 + bg1)          E: T1 <- r * h1                       (3/4 of the u half under it)
//   C1 : c = tanh([T0 | T1] Wc1 + bc1)                 E: h1 <- m*h1' ; H1 <- h1             (last quarter of u behind C1)
// The issuer is in-order and so is the tensor pipe: a product that is NOT on the recurrence's critical path (u, sbias) is placed
// where an epilogue runs anyway, and a streamed tile is never waited for in front of a product the recurrence needs.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "sm100.cuh"

namespace htcn {
using namespace sm100;

int32_t k3_prepare_rep(const float* const* gate_w, const float* const* gate_b, const float* const* cand_w,
                       const float* const* cand_b, const float* w_in_state, float* scratch, cudaStream_t st,
                       const uint8_t** w_out, float** bias_out);

namespace k3t {
constexpr int kSub = 128 * 64 * 2;                // one weight sub-tile: 128 hidden units x 64 k, bf16, 128-byte swizzle = 16 KB
constexpr int kSubPerStep = 28;                   // G0r 4 | G0u 4 | SB 4 | C0 4 | G1r 4 | G1u 4 | C1 4
constexpr int kUsersPerThread = 16;
enum Slot { kX = 0, kH0 = 1, kT0 = 2, kH1 = 3, kT1 = 4, kSlots = 5 };
enum Product { kG0r = 0, kG0u, kSB, kC0, kG1r, kG1u, kC1 };

template <int kNU, int kRes, int kStages, bool kTs = false>
struct Cfg {
  static constexpr int kEpiWarps = 4 * (kNU / kUsersPerThread);     // 4 TMEM lane quarters x user groups of 16
  static constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;
  static constexpr int kThreads = 32 * (kEpiWarps + 2);
  // K chunks (8 k = 16 B per user row) are kNU*16 + 16 bytes apart: the 32 lanes of a warp store 2 bytes each into 4
  // consecutive chunks of one user row, the 16 bytes of padding put those on different banks
  static constexpr int kChunkStride = kNU * 16 + 16;
  static constexpr int kSlotBytes = 16 * kChunkStride;
  static constexpr uint32_t kColR = 0, kColU = kNU, kColC = 2 * kNU, kColSb = 3 * kNU;
  // kTs: three weight tiles (K = 256 -> 128 columns of bf16 pairs each) live in tensor memory behind the accumulators
  static constexpr uint32_t kColW = 128, kTmemCols = kTs ? 512 : 4 * kNU;
  static_assert(!kTs || 4 * kNU <= 128, "accumulators must fit in front of the weight columns");
  struct alignas(1024) Smem {
    uint8_t res[kRes > 0 ? kRes : 1][kSub];
    uint8_t ring[kStages][kSub];
    alignas(16) uint8_t act[kSlots][kSlotBytes];
    uint64_t w_full[kStages], w_empty[kStages], res_full, ts_ready, act_ready, acc_r, acc_u, acc_c, acc_sb;
    uint32_t tmem_base;
  };
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
// first weight tile (of the 14 [128 n][128 k] tiles k3_prepare_weights writes) of a product
__host__ __device__ constexpr int tile_of(int product) {
  return product == kSB ? 0 : product == kG0r ? 2 : product == kG0u ? 4 : product == kC0 ? 6 : product == kG1r ? 8
         : product == kG1u ? 10 : 12;
}
// Where a product's weight tile (K = 256: four sub-tiles) comes from.  Plain plan: the first kRes sub-tiles of the step are
// resident in shared memory, the rest stream through the ring.  kTs plan: the products every step WAITS for -- G0r, C0, G1r --
// read their A operand from TENSOR MEMORY (tcgen05.mma with A in TMEM: no shared-memory read of the 128-row operand, 16 clk
// per MMA instead of 64, and 192 KB of weights that neither occupy shared memory nor cross the L2->SM path again); G0u and C1
// are resident in shared memory; only SB and G1u (128 KB per step instead of 320) stream.
enum Src { kRing = 0, kResident = 1, kTmem = 2 };
__host__ __device__ constexpr int src_of(int p, bool ts, int n_res) {
  return ts ? ((p == kG0r || p == kC0 || p == kG1r) ? kTmem : (p == kG0u || p == kC1) ? kResident : kRing)
            : (p * 4 < n_res ? kResident : kRing);
}
__host__ __device__ constexpr int res_first(int p, bool ts) { return ts ? (p == kG0u ? 0 : 4) : p * 4; }
__host__ __device__ constexpr int ts_index(int p) { return p == kG0r ? 0 : p == kC0 ? 1 : 2; }
}  // namespace k3t

template <int kNU, int kRes, int kStages, bool kTs, bool kGates>
__global__ void __launch_bounds__(k3t::Cfg<kNU, kRes, kStages, kTs>::kThreads, 1)
k3_gru_bf16_t(const uint8_t* __restrict__ w_img, const float* __restrict__ yp, const float* __restrict__ mask,
              const float* __restrict__ state_in, const float* __restrict__ bias_all /* [bg0 256][bc0 128][bg1 256][bc1 128] */,
              int B, int S, int do_sbias, float* __restrict__ state_pre, float* __restrict__ sbias,
              float* __restrict__ state_out,
              float* __restrict__ gates_save /* kGates: [S][2][3][B][128] = r, u, c of every cell call (for htcn_gru_backward) */) {
  using namespace k3t;
  using C = Cfg<kNU, kRes, kStages, kTs>;
  using Smem = typename C::Smem;
  extern __shared__ uint8_t smem_raw[];
  auto& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // -DHTCN_K3_TRACE (HTCN_NVCC_EXTRA of hiertcn_b200/build.py): CTA 1 records clock64 at the phase boundaries of step 3 -- MMA
  // warp events 0..10, epilogue warp 0 events (layer * 10 + 0..8) -- and prints them relative to the step's first event
#ifdef HTCN_K3_TRACE
  __shared__ uint32_t k3_trace[2][24];
#define K3T_TR(w, k) do { if (blockIdx.x == 1 && s == 3 && lane == 0) k3_trace[w][k] = (uint32_t)clock64(); } while (0)
#else
#define K3T_TR(w, k) do { } while (0)
#endif

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.w_full[s], 1);
      mbar_init(&sm.w_empty[s], 1);
    }
    mbar_init(&sm.res_full, 1);
    mbar_init(&sm.ts_ready, C::kEpiWarps);
    mbar_init(&sm.act_ready, C::kEpiWarps);                       // one arrive per epilogue warp
    mbar_init(&sm.acc_r, 1);
    mbar_init(&sm.acc_u, 1);
    mbar_init(&sm.acc_c, 1);
    mbar_init(&sm.acc_sb, 1);
    fence_barrier_init();
  }
  if (warp == C::kMmaWarp) tmem_alloc<C::kTmemCols>(&sm.tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, sm.tmem_base, 0);

  if (warp == C::kProducerWarp) {
    // ===================== weight producer: resident sub-tiles once, the rest of every step through the ring ==========
    if (lane == 0) {
      if (kRes > 0) {
        mbar_arrive_expect_tx(&sm.res_full, kRes * kSub);
        for (int p = 0; p < 7; ++p)
          if (src_of(p, kTs, kRes) == kResident)
            for (int kt = 0; kt < 4; ++kt)
              bulk_load_1d(sm.res[res_first(p, kTs) + kt], w_img + (size_t)((tile_of(p) + (kt >> 1)) * 2 + (kt & 1)) * kSub, kSub,
                           &sm.res_full);
      }
      // the ring follows the MMA issuer's order: G0r G0u | C0 SB | G1r G1u[0..2] | C1 G1u[3]
      constexpr int seg_p[8] = {kG0r, kG0u, kC0, kSB, kG1r, kG1u, kC1, kG1u};
      constexpr int seg_k0[8] = {0, 0, 0, 0, 0, 0, 0, 3}, seg_k1[8] = {4, 4, 4, 4, 4, 3, 4, 4};
      long long n = 0;
      for (int s = 0; s < S; ++s) {
        for (int g = 0; g < 8; ++g) {
          const int p = seg_p[g];
          if (src_of(p, kTs, kRes) != kRing || (p == kSB && !do_sbias)) continue;
          for (int kt = seg_k0[g]; kt < seg_k1[g]; ++kt) {
            const int st = (int)(n % kStages);
            mbar_wait_relaxed(&sm.w_empty[st], (uint32_t)(((n / kStages) & 1) ^ 1));
            mbar_arrive_expect_tx(&sm.w_full[st], kSub);
            bulk_load_1d(sm.ring[st], w_img + (size_t)((tile_of(p) + (kt >> 1)) * 2 + (kt & 1)) * kSub, kSub, &sm.w_full[st]);
            ++n;
          }
        }
      }
    }
  } else if (warp == C::kMmaWarp) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(128, kNU);
    const uint32_t act0 = smem_u32(sm.act);
    long long n = 0, n_act = 0;
    // one product = 4 sub-tiles (K = 256 = two operand slots) into the kNU accumulator columns at d_col
    auto product = [&](auto pc, uint32_t d_col, int slot_a, int slot_b, int kt0 = 0, int kt1 = 4) {
      constexpr int p = decltype(pc)::value;
      constexpr int src = src_of(p, kTs, kRes);
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        if (kt < kt0 || kt >= kt1) continue;
        uint32_t w_base = 0;
        int st = 0;
        if (src == kResident) {
          w_base = smem_u32(sm.res[res_first(p, kTs) + kt]);
        } else if (src == kRing) {
          st = (int)(n % kStages);
          mbar_wait(&sm.w_full[st], (uint32_t)((n / kStages) & 1));
          tc_fence_after_sync();
          w_base = smem_u32(sm.ring[st]);
        }
        const uint32_t slot = act0 + (uint32_t)((kt >> 1) ? slot_b : slot_a) * C::kSlotBytes;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t db = make_desc_nosw(slot + (uint32_t)((kt & 1) * 8 + kk * 2) * C::kChunkStride, C::kChunkStride);
          if (src == kTmem) {           // A = 8 columns (16 k as bf16 pairs) of the tile's 128 weight columns
            if (leader) umma_bf16_ts(tmem + d_col, tmem + C::kColW + ts_index(p) * 128 + (kt * 4 + kk) * 8, db, idesc, !(kt == 0 && kk == 0));
          } else {
            const uint64_t da = make_desc_k_sw128(w_base + kk * 32);
            if (leader) umma_bf16(tmem + d_col, da, db, idesc, !(kt == 0 && kk == 0));
          }
        }
        if (src == kRing) {
          if (leader) umma_commit(&sm.w_empty[st]);
          ++n;
        }
      }
    };
    auto wait_operand = [&]() {
      mbar_wait(&sm.act_ready, (uint32_t)(n_act & 1));
      tc_fence_after_sync();
      ++n_act;
    };
    if (kRes > 0) {
      mbar_wait(&sm.res_full, 0);
      tc_fence_after_sync();
    }
    if (kTs) {
      mbar_wait(&sm.ts_ready, 0);                                 // the epilogue warps have stored the TMEM-resident tiles
      tc_fence_after_sync();
    }
    for (int s = 0; s < S; ++s) {
      wait_operand();                                             // X, H0, H1 of this step
      K3T_TR(0, 0);
      product(std::integral_constant<int, kG0r>{}, C::kColR, kX, kH0);
      if (leader) umma_commit(&sm.acc_r);
      K3T_TR(0, 1);
      product(std::integral_constant<int, kG0u>{}, C::kColU, kX, kH0);             // runs under E_g0
      if (leader) umma_commit(&sm.acc_u);
      K3T_TR(0, 2);
      K3T_TR(0, 3);
      wait_operand();                                             // T0 = r * h0
      K3T_TR(0, 4);
      product(std::integral_constant<int, kC0>{}, C::kColC, kX, kT0);
      if (leader) umma_commit(&sm.acc_c);
      if (do_sbias) {                                             // runs under E_c0 (which rewrites H0 only after acc_sb)
        product(std::integral_constant<int, kSB>{}, C::kColSb, kH0, kH1);
        if (leader) umma_commit(&sm.acc_sb);
      }
      K3T_TR(0, 5);
      wait_operand();                                             // T0 = h0' (unmasked)
      K3T_TR(0, 6);
      product(std::integral_constant<int, kG1r>{}, C::kColR, kT0, kH1);
      if (leader) umma_commit(&sm.acc_r);
      K3T_TR(0, 7);
      product(std::integral_constant<int, kG1u>{}, C::kColU, kT0, kH1, 0, 3);      // three of its four sub-tiles under E_g1 ...
      K3T_TR(0, 8);
      wait_operand();                                             // T1 = r * h1 (and H0 = m * h0', X = x(s+1))
      K3T_TR(0, 9);
      product(std::integral_constant<int, kC1>{}, C::kColC, kT0, kT1);
      if (leader) umma_commit(&sm.acc_c);
      product(std::integral_constant<int, kG1u>{}, C::kColU, kT0, kH1, 3, 4);      // ... the last one behind the candidate: a
      if (leader) umma_commit(&sm.acc_u);                         // streamed tile must never make the issuer wait in front of a
      K3T_TR(0, 10);                                              // product the recurrence is waiting for
    }
  } else {
    // ===================== epilogue: thread = hidden unit j of 16 users =====================
    const int quarter = warp & 3, ug = warp >> 2;
    const int j = quarter * 32 + lane;                            // hidden unit = TMEM lane
    const int u0 = ug * kUsersPerThread;                          // first user (row of the operand slots) of this thread
    const long long b0 = (long long)blockIdx.x * kNU + u0;
    const int n_ok = (int)max(0LL, min((long long)kUsersPerThread, (long long)B - b0));
    const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)u0;
    // shared address of (user u0, k = j) in slot 0; users are 16 B apart, slots kSlotBytes apart
    const uint32_t a0 = smem_u32(sm.act) + (uint32_t)(j >> 3) * C::kChunkStride + (uint32_t)u0 * 16 + (uint32_t)(j & 7) * 2;
    float bgr[2], bgu[2], bcc[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      bgr[l] = __ldg(bias_all + l * 384 + j);
      bgu[l] = __ldg(bias_all + l * 384 + 128 + j);
      bcc[l] = __ldg(bias_all + l * 384 + 256 + j);
    }
    float h[2][kUsersPerThread], xn[kUsersPerThread], m[kUsersPerThread], u[kUsersPerThread];
    uint32_t par_r = 0, par_u = 0, par_c = 0, par_sb = 0;
    auto put = [&](int slot, int i, float v) { sts_bf16(a0 + (uint32_t)slot * C::kSlotBytes + (uint32_t)i * 16, v); };
    // every lane orders its own operand stores before the async proxy, then ONE lane per warp arrives (256 arrives on one
    // mbarrier serialise)
    auto publish = [&]() {
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.act_ready);
    };
    auto wait_acc = [&](uint64_t* bar, uint32_t& par) {
      mbar_wait(bar, par);
      par ^= 1;
      tc_fence_after_sync();
    };
    auto ld_acc = [&](uint32_t col, float (&v)[kUsersPerThread]) {
      uint32_t r[16];
      tmem_ld_32x16(t_lane + col, r);
      tmem_ld_wait(r);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
    };
    if (kTs) {
      // weight tiles of G0r, C0, G1r -> tensor memory: lane = hidden unit j = row of the tile, column c of the tile holds
      // (k = 2c, 2c+1) as a bf16 pair.  This thread copies row j of two of the tile's four 64-k sub-tile images (ug picks the
      // k half): 8 swizzled 16-byte chunks each -> 32 columns, one tcgen05.st
      const uint32_t t_w = tmem + ((uint32_t)(quarter * 32) << 16) + C::kColW;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int p = t == 0 ? kG0r : t == 1 ? kC0 : kG1r;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int sub = ug * 2 + h2;                             // 64-k sub-tile of the K = 256 tile
          const uint4* row = reinterpret_cast<const uint4*>(w_img + (size_t)((tile_of(p) + (sub >> 1)) * 2 + (sub & 1)) * kSub +
                                                            (size_t)j * 128);
          uint32_t v[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 q = __ldg(row + (c ^ (j & 7)));
            v[c * 4] = q.x; v[c * 4 + 1] = q.y; v[c * 4 + 2] = q.z; v[c * 4 + 3] = q.w;
          }
          tmem_st_32x32(t_w + t * 128 + sub * 32, v);
        }
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.ts_ready);
    }
    // prologue: state and the first input -> registers and operand slots
#pragma unroll
    for (int i = 0; i < kUsersPerThread; ++i) {
      const bool ok = i < n_ok;
      h[0][i] = ok ? __ldg(state_in + (b0 + i) * 256 + j) : 0.f;
      h[1][i] = ok ? __ldg(state_in + (b0 + i) * 256 + 128 + j) : 0.f;
      xn[i] = ok ? __ldg(yp + (b0 + i) * kDim + j) : 0.f;
      put(kH0, i, h[0][i]);
      put(kH1, i, h[1][i]);
      put(kX, i, xn[i]);
    }
    publish();
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        const int slot_t = l == 0 ? kT0 : kT1;
        float v[kUsersPerThread];
        if (warp == 0) K3T_TR(1, l * 10 + 0);
        wait_acc(&sm.acc_r, par_r);                                // ---- E_g: T <- r * h
        if (warp == 0) K3T_TR(1, l * 10 + 1);
        ld_acc(C::kColR, v);
        if (warp == 0) K3T_TR(1, l * 10 + 2);
#pragma unroll
        for (int i = 0; i < kUsersPerThread; ++i) {
          v[i] = sigmoid_fast(v[i] + bgr[l]);
          put(slot_t, i, v[i] * h[l][i]);
        }
        publish();
        if (warp == 0) K3T_TR(1, l * 10 + 3);
        // training: the gate activations go out behind the publish, under the product the issuer has just been released for
        float* gs = kGates ? gates_save + ((long long)(s * 2 + l) * 3 * B + b0) * kDim + j : nullptr;
        if (kGates) {
#pragma unroll
          for (int i = 0; i < kUsersPerThread; ++i)
            if (i < n_ok) gs[(long long)i * kDim] = v[i];                                  // r
        }
        if (l == 0) {
          // ---- under the candidate product: emit the state before the step, fetch the mask and the next input.  NOT at the top
          // of the step: a fence.proxy.async (every publish) waits for the thread's outstanding global loads, and these take
          // ~1500 clk from HBM -- placed before the first publish they sat on the critical path of every step
          if (state_pre) {
#pragma unroll
            for (int i = 0; i < kUsersPerThread; ++i)
              if (i < n_ok) {
                float* o = state_pre + ((long long)s * B + b0 + i) * 256 + j;
                o[0] = h[0][i];
                o[128] = h[1][i];
              }
          }
#pragma unroll
          for (int i = 0; i < kUsersPerThread; ++i) {
            const bool ok = i < n_ok;
            m[i] = ok ? __ldg(mask + (long long)s * B + b0 + i) : 0.f;
            xn[i] = (ok && s + 1 < S) ? __ldg(yp + ((long long)(s + 1) * B + b0 + i) * kDim + j) : 0.f;
          }
        }
        if (l == 0) {                                              // layer 0: u arrives under E_g0, the candidate after it
          wait_acc(&sm.acc_u, par_u);
          if (warp == 0) K3T_TR(1, l * 10 + 4);
          ld_acc(C::kColU, v);
#pragma unroll
          for (int i = 0; i < kUsersPerThread; ++i) u[i] = sigmoid_fast(v[i] + bgu[l]);
        }
        if (warp == 0) K3T_TR(1, l * 10 + 5);
        wait_acc(&sm.acc_c, par_c);                                // ---- E_c: h' = u*h + (1-u)*c
        if (warp == 0) K3T_TR(1, l * 10 + 6);
        float c[kUsersPerThread];
        ld_acc(C::kColC, c);
#pragma unroll
        for (int i = 0; i < kUsersPerThread; ++i) c[i] = tanh_fast(c[i] + bcc[l]);
        if (l == 1) {                                              // layer 1: the last sub-tile of u is issued behind the candidate
          wait_acc(&sm.acc_u, par_u);
          if (warp == 0) K3T_TR(1, l * 10 + 4);
          ld_acc(C::kColU, v);
#pragma unroll
          for (int i = 0; i < kUsersPerThread; ++i) u[i] = sigmoid_fast(v[i] + bgu[l]);
        }
        if (warp == 0) K3T_TR(1, l * 10 + 7);
#pragma unroll
        for (int i = 0; i < kUsersPerThread; ++i) {
          const float o = fmaf(u[i], h[l][i] - c[i], c[i]);        // UNMASKED: the input of the layer above
          h[l][i] = m[i] * o;                                      // state *= mask (model_hier.py:93)
          if (l == 0) put(kT0, i, o);
          else put(kH1, i, h[1][i]);
        }
        if (l == 0 || s + 1 < S) publish();
        if (warp == 0) K3T_TR(1, l * 10 + 8);
        if (kGates) {
#pragma unroll
          for (int i = 0; i < kUsersPerThread; ++i)
            if (i < n_ok) {
              gs[((long long)B + i) * kDim] = u[i];                                        // u
              gs[(2LL * B + i) * kDim] = c[i];                                             // c
            }
        }
        if (l == 0) {
          // under G1r: sbias[s] out (its product ran under this epilogue), THEN the new masked h0 -- the sbias product reads the
          // H0 slot -- and the next step's input (X is free since C0); all three are published by the next publish (E_g1)
          if (do_sbias) {
            wait_acc(&sm.acc_sb, par_sb);
            ld_acc(C::kColSb, v);
#pragma unroll
            for (int i = 0; i < kUsersPerThread; ++i)
              if (i < n_ok) sbias[((long long)s * B + b0 + i) * kDim + j] = v[i];
          }
#pragma unroll
          for (int i = 0; i < kUsersPerThread; ++i) {
            put(kH0, i, h[0][i]);
            put(kX, i, xn[i]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kUsersPerThread; ++i)
      if (i < n_ok) {
        float* o = state_out + (b0 + i) * 256 + j;
        o[0] = h[0][i];
        o[128] = h[1][i];
      }
  }
  tc_fence_before_sync();
  __syncthreads();
#ifdef HTCN_K3_TRACE
  if (blockIdx.x == 1 && tid == 0 && S > 3) {
    for (int k = 0; k <= 10; ++k) printf("k3t mma %2d: %6u\n", k, k3_trace[0][k] - k3_trace[0][0]);
    for (int k = 0; k < 19; ++k) printf("k3t epi %2d: %6u\n", k, k3_trace[1][k] - k3_trace[0][0]);
  }
#endif
#undef K3T_TR
  if (warp == C::kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc<C::kTmemCols>(tmem);
  }
}

template <int kNU, int kRes, int kStages, bool kTs, bool kGates = false>
static int32_t launch_t(const uint8_t* tw, const float* yp, const float* mask, const float* state_in, const float* bias_dev,
                        int B, int S, float* state_pre, float* sbias, float* state_out, cudaStream_t st,
                        float* gates_save = nullptr) {
  using C = k3t::Cfg<kNU, kRes, kStages, kTs>;
  const size_t smem = sizeof(typename C::Smem) + 1024;
  static_assert(sizeof(typename C::Smem) + 1024 <= 232448, "shared-memory budget of one CTA");
  auto kern = k3_gru_bf16_t<kNU, kRes, kStages, kTs, kGates>;
  HTCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<ceil_div(B, kNU), C::kThreads, smem, st>>>(tw, yp, mask, state_in, bias_dev, B, S, sbias != nullptr, state_pre,
                                                   sbias, state_out, gates_save);
  HTCN_LAUNCH_CHECK("k3_gru_bf16_t");
  return HTCN_OK;
}

// variant 0: layer-0 gates resident (128 KB), 3-stage ring; 1: only their r half resident, 7-stage ring; 2: three weight
// tiles in tensor memory.  (64 users per CTA -- kNU = 64, 16 epilogue warps -- was tried: 108 us against 84, the per-SM
// weight stream is the same and 576 threads spill.)  gates_save != NULL: the training forward (variant 2 + the gate
// activations of every cell call written out for htcn_gru_backward).
int32_t gru_sessions_bf16_t(const float* yp, const float* mask, const float* state_in, const float* const* gate_w,
                            const float* const* gate_b, const float* const* cand_w, const float* const* cand_b,
                            const float* w_in_state, int B, int S, float* state_pre, float* sbias, float* state_out,
                            float* scratch, int variant, cudaStream_t st, float* gates_save) {
  float* bias_dev;
  const uint8_t* tw;
  int32_t rc = k3_prepare_rep(gate_w, gate_b, cand_w, cand_b, w_in_state, scratch, st, &tw, &bias_dev);
  if (rc) return rc;
  if (gates_save)
    return launch_t<32, 8, 3, true, true>(tw, yp, mask, state_in, bias_dev, B, S, state_pre, sbias, state_out, st, gates_save);
  if (variant == 2) return launch_t<32, 8, 3, true>(tw, yp, mask, state_in, bias_dev, B, S, state_pre, sbias, state_out, st);
  if (variant == 1) return launch_t<32, 4, 7, false>(tw, yp, mask, state_in, bias_dev, B, S, state_pre, sbias, state_out, st);
  return launch_t<32, 8, 3, false>(tw, yp, mask, state_in, bias_dev, B, S, state_pre, sbias, state_out, st);
}

}  // namespace htcn
