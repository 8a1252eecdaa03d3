// K3 (bf16 tier), cluster variant: GRU over sessions with the recurrent weights RESIDENT in shared memory.
//
// A cluster of 4 CTAs owns 128 users (MMA M = 128) and runs all S steps.  CTA c of the cluster owns the hidden
// columns [32c, 32c+32) of every product: its slice of ALL weights of the step -- W_in[D:] (sbias), the gate and
// candidate kernels of both layers, 112 KB of bf16 -- is loaded ONCE (cp.async.bulk) and never leaves shared memory,
// 128 SMs work on a batch of 4096 users instead of 32, and each of the five dependent GEMM phases of a step is an
// N = 64 / N = 32 product (16 tcgen05.mma) instead of N = 256 / 128.
//
// Every CTA keeps a full copy of the GEMM operand, three 128-column slots [x | h0 | h1] (128 users x 384 bf16, no-swizzle
// K-major, 96 KB); a product reads two adjacent slots: [x | h0] for layer 0, [h0 | h1] for layer 1 and for sbias, so the
// sbias product rides in the same phase as the gates of layer 0 and a step is FOUR dependent phases.
// After a phase, an epilogue thread holds 8 columns of one user (r*h, or h'): it writes that ONE 16-byte K chunk into the
// operand buffers of all 4 CTAs through distributed shared memory (st.async / st.shared::cluster), so the h' slices are
// exchanged without touching L2.  The fp32 recurrent state slice (8 + 8 values per thread) lives in registers.
//
// Synchronisation per phase (two mbarriers per CTA, no cluster-wide barrier.cluster on the critical path):
//   act_ready : 4 CTAs x 16 epilogue warps arrive remotely (fence.proxy.async by every writer, then one
//               fence.acq_rel.cluster + 4 relaxed remote arrives per warp) -> the MMA warp of this CTA issues the phase;
//   acc_ready : every CTA's tcgen05.commit is MULTICAST to the acc_ready barriers of all 4 CTAs (count 4): when it
//               completes, this CTA's accumulator is ready AND all four operand buffers are free to be rewritten.
//
// Step s (customed_gru_cell.py:309-337 per layer, :1050-1073 stacking; model_hier.py:54-55,91,93):
//   P1 : [r|u] = sigmoid([x | h0] Wg0 + bg0),  sbias[s] = [h0 | h1] @ W_in[D:]      E: h0 slot <- r * h0
//   P2 : c = tanh([x | r*h0] Wc0 + bc0)        E: h0' = u*h0 + (1-u)*c ; state <- m*h0' ; h0 slot <- h0'
//   P3 : [r|u] = sigmoid([h0' | h1] Wg1 + bg1)                                       E: h1 slot <- r * h1
//   P4 : c = tanh([h0' | r*h1] Wc1 + bc1)      E: state <- m*h1' ; slots <- [x_{s+1} | m*h0' | m*h1']
#include <cstdlib>

#include "common.cuh"
#include "sm100.cuh"

namespace htcn {
using namespace sm100;

namespace k3c {
constexpr int kM = 128;                           // users per cluster
constexpr int kCl = 4;                            // CTAs per cluster
constexpr int kSlice = 128 / kCl;                 // hidden columns per CTA
constexpr int kSlotX = 0, kSlotH0 = 16, kSlotH1 = 32;   // operand slots (first 16-byte K chunk): [x | h0 | h1], 128 columns each
constexpr int kActBytes = 48 * kM * 16;           // 48 x 16-byte K chunks x 128 rows = 96 KB; a product reads 32 consecutive chunks
constexpr int kEpiWarps = 16;                     // 4 TMEM lane quarters x 4 groups of 8 columns
constexpr int kThreads = 32 * (kEpiWarps + 1);    // warps 0-15 epilogue, 16 = weight load + MMA issuer
constexpr int kMmaWarp = kEpiWarps;
// per-CTA weight blob, every matrix [32 K chunks][n rows][8 bf16]:  sbias 32 rows | gates0 64 | cand0 32 | gates1 64 | cand1 32
constexpr int kOffSb = 0, kOffG0 = 32, kOffC0 = 96, kOffG1 = 128, kOffC1 = 192, kRowsTotal = 224;
constexpr int kBlobBytes = kRowsTotal * 256 * 2;  // 112 KB
// TMEM columns: [r 32 | u 32] gates, [cand 32], [sbias 32]
constexpr uint32_t kColG = 0, kColC = 64, kColSb = 96, kTmemCols = 128;

struct alignas(1024) Smem {
  uint8_t w[kBlobBytes];                          // 112 KB, resident
  uint8_t act[kActBytes];                         // 96 KB
  float bg[2][256];
  float bc[2][128];
  uint64_t w_full, acc_ready, act_ready;
  uint32_t tmem_base;
};

// no-swizzle K-major operand: 8-row x 16-byte core matrices, 8-row groups 128 B apart, K chunks `lbo` bytes apart
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t smem_addr, uint32_t lbo) {
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

__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 16-byte store into a peer CTA's shared memory that completes `16` transaction bytes on the peer's mbarrier: no
// cluster-scope fence (MEMBAR.GPU) is needed to publish it
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // bounded, like mbar_wait
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("htcn: k3 cluster mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
// all MMAs issued so far by this thread -> arrive on `bar` (same smem offset) in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
// tmem_ld_32x8: sm100.cuh
__device__ __forceinline__ void tmem_ld_wait8(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
}  // namespace k3c

// kAsync: the operand chunks travel as st.async (complete_tx on the destination's act_ready barrier; the barrier counts
// this CTA's 16 epilogue warps + the phase's transaction bytes); otherwise as plain st.shared::cluster published by a
// cluster-scope fence and 4 remote arrives per warp.
template <bool kAsync>
__global__ void __cluster_dims__(k3c::kCl, 1, 1) __launch_bounds__(k3c::kThreads, 1)
k3_gru_bf16_cluster(const uint8_t* __restrict__ w_blob /* [4][kBlobBytes] */, const float* __restrict__ yp,
                    const float* __restrict__ mask, const float* __restrict__ state_in,
                    const float* __restrict__ bias_all /* [bg0 256][bc0 128][bg1 256][bc1 128] */, int B, int S, int do_sbias,
                    float* __restrict__ state_pre, float* __restrict__ sbias, float* __restrict__ state_out) {
  using namespace k3c;
  extern __shared__ uint8_t smem_raw[];
  auto& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();
  const int tile = blockIdx.x / kCl;

  if (tid == 0) {
    mbar_init(&sm.w_full, 1);
    mbar_init(&sm.acc_ready, kCl);
    mbar_init(&sm.act_ready, kAsync ? kEpiWarps : kCl * kEpiWarps);
    fence_barrier_init();
  }
  for (int i = tid; i < 768; i += kThreads) {
    const int l = i / 384, j = i % 384;
    if (j < 256) sm.bg[l][j] = bias_all[i];
    else sm.bc[l][j - 256] = bias_all[i];
  }
  if (warp == kMmaWarp) tmem_alloc<kTmemCols>(&sm.tmem_base);
  tc_fence_before_sync();
  cluster_sync_all();                              // barriers of all 4 CTAs are initialised before any remote arrive
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, sm.tmem_base, 0);

  if (warp == kMmaWarp) {
    // ===================== resident weights + MMA issuer =====================
    const bool leader = elect_one();
    if (leader) {
      mbar_arrive_expect_tx(&sm.w_full, kBlobBytes);
      const uint8_t* src = w_blob + (size_t)crank * kBlobBytes;
      for (int o = 0; o < kBlobBytes; o += 16384) bulk_load_1d(sm.w + o, src + o, 16384, &sm.w_full);
    }
    mbar_wait(&sm.w_full, 0);
    const uint32_t act0 = smem_u32(sm.act), w0 = smem_u32(sm.w);
    uint32_t n_act = 0;
    auto begin_phase = [&]() {
      mbar_wait_cluster(&sm.act_ready, n_act & 1);
      ++n_act;
      if (kAsync) fence_proxy_async_all();                       // st.async data (generic proxy) -> UMMA operand reads (async proxy)
      tc_fence_after_sync();
    };
    // D[128 x n_rows] = operand chunks [a_chunk, a_chunk + 32) (K = 256) * W[n_rows x 256]^T
    auto product = [&](uint32_t d_col, int a_chunk, int row_off, int n_rows, uint32_t idesc) {
      const uint32_t wb = w0 + (uint32_t)row_off * 512;           // matrices are stored one after the other: 512 B per row
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint64_t da = make_desc_nosw(act0 + (uint32_t)(a_chunk + 2 * k) * (kM * 16), kM * 16);
        const uint64_t db = make_desc_nosw(wb + (uint32_t)(2 * k) * (uint32_t)(n_rows * 16), (uint32_t)(n_rows * 16));
        if (leader) umma_bf16(tmem + d_col, da, db, idesc, k > 0);
      }
    };
    auto end_phase = [&]() {
      if (leader) umma_commit_multicast(&sm.acc_ready, (uint16_t)((1u << kCl) - 1));
    };
    constexpr uint32_t idesc64 = make_idesc_bf16(kM, 64), idesc32 = make_idesc_bf16(kM, 32);
    for (int s = 0; s < S; ++s) {
      begin_phase();                                              // P1: operand [x | h0 | h1]
      product(kColG, kSlotX, kOffG0, 64, idesc64);                //   gates of layer 0 from [x | h0]
      if (do_sbias) product(kColSb, kSlotH0, kOffSb, 32, idesc32);//   sbias[s] from [h0 | h1], same phase
      end_phase();
      begin_phase();                                              // P2: [x | r*h0]
      product(kColC, kSlotX, kOffC0, 32, idesc32);
      end_phase();
      begin_phase();                                              // P3: [h0' | h1]
      product(kColG, kSlotH0, kOffG1, 64, idesc64);
      end_phase();
      begin_phase();                                              // P4: [h0' | r*h1]
      product(kColC, kSlotH0, kOffC1, 32, idesc32);
      end_phase();
    }
  } else {
    // ===================== epilogue: thread = 8 hidden columns of one user =====================
    const int quarter = warp & 3, sub = warp >> 2;
    const int r = quarter * 32 + lane;
    const int b = tile * kM + r;
    const bool ok = b < B;
    const int col = (int)crank * kSlice + sub * 8;                // first of this thread's 8 hidden columns
    const int chunk = col >> 3;                                   // its 16-byte K chunk inside an operand slot
    const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16);
    uint8_t* act_row = sm.act + r * 16;
    uint32_t act_remote[kCl], bar_remote[kCl];
#pragma unroll
    for (int d = 0; d < kCl; ++d) {
      act_remote[d] = mapa_u32(smem_u32(act_row), (uint32_t)d);
      bar_remote[d] = mapa_u32(smem_u32(&sm.act_ready), (uint32_t)d);
    }
    // this thread's chunk of operand slot `slot0` (kSlotH0 / kSlotH1) in all 4 CTAs <- bf16(v)
    auto put_all = [&](int slot0, const float (&v)[8]) {
      const uint4 p = pack8(v);
      const uint32_t off = (uint32_t)(slot0 + chunk) * (kM * 16);
#pragma unroll
      for (int d = 0; d < kCl; ++d) {
        if (kAsync) st_async_v4(act_remote[d] + off, p, bar_remote[d]);
        else st_cluster_v4(act_remote[d] + off, p);
      }
    };
    // the x slot is written locally: every CTA stages the full input of its 128 users (this thread: 32 columns).  The
    // rows of step s+1 are fetched (and packed to bf16: 16 registers) at the top of step s, a whole step before their use:
    // loading them where they are stored put ~1 us of HBM latency on the critical path of every step (ncu: long_scoreboard)
    uint4 xp[4];
    auto fetch_x = [&](int s) {
      const float4* x = reinterpret_cast<const float4*>(yp + ((long long)s * B + b) * kDim + sub * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 q0 = ok ? __ldg(x + 2 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 q1 = ok ? __ldg(x + 2 * i + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        xp[i] = pack8(v);
      }
    };
    auto put_x = [&]() {
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(act_row + (kSlotX + sub * 4 + i) * (kM * 16)) = xp[i];
    };
    // `slots` = operand slots every thread of the cluster wrote with put_all in this phase (32 KB each per CTA buffer)
    auto signal = [&](int slots) {
      fence_proxy_async_all();                                    // generic-proxy writes (local + remote) -> async proxy
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (kAsync) {
          if (warp == 0) mbar_arrive_expect_tx(&sm.act_ready, (uint32_t)slots * (kM * 128 * 2));
          else mbar_arrive(&sm.act_ready);
        } else {
          fence_acq_rel_cluster();
#pragma unroll
          for (int d = 0; d < kCl; ++d) mbar_arrive_remote_relaxed(bar_remote[d]);
        }
      }
    };
    uint32_t n_acc = 0;
    auto wait_acc = [&]() {
      mbar_wait(&sm.acc_ready, n_acc & 1);
      ++n_acc;
      tc_fence_after_sync();
    };
    float h[2][8];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      const float4 a0 = ok ? __ldg(reinterpret_cast<const float4*>(state_in + (long long)b * 256 + l * 128 + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 a1 = ok ? __ldg(reinterpret_cast<const float4*>(state_in + (long long)b * 256 + l * 128 + col) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      h[l][0] = a0.x; h[l][1] = a0.y; h[l][2] = a0.z; h[l][3] = a0.w; h[l][4] = a1.x; h[l][5] = a1.y; h[l][6] = a1.z; h[l][7] = a1.w;
    }
    fetch_x(0);
    put_x();                                                      // operand of the first phase: [x_0 | h0 | h1]
    put_all(kSlotH0, h[0]);
    put_all(kSlotH1, h[1]);
    signal(2);
    for (int s = 0; s < S; ++s) {
      const float m = ok ? __ldg(mask + (long long)s * B + b) : 0.f;
      if (s + 1 < S) fetch_x(s + 1);
      if (state_pre && ok) {
#pragma unroll
        for (int l = 0; l < 2; ++l) {
          float4* o = reinterpret_cast<float4*>(state_pre + ((long long)s * B + b) * 256 + l * 128 + col);
          o[0] = make_float4(h[l][0], h[l][1], h[l][2], h[l][3]);
          o[1] = make_float4(h[l][4], h[l][5], h[l][6], h[l][7]);
        }
      }
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        wait_acc();                                               // ---- E_g: hidden slot of layer l <- r * h
        uint32_t vr[8], vu[8];
        tmem_ld_32x8(t_lane + kColG + sub * 8, vr);
        tmem_ld_32x8(t_lane + kColG + 32 + sub * 8, vu);
        tmem_ld_wait8(vr);
        tmem_ld_wait8(vu);
        float rh[8], u[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          rh[e] = sigmoid_fast(__uint_as_float(vr[e]) + sm.bg[l][col + e]) * h[l][e];
          u[e] = sigmoid_fast(__uint_as_float(vu[e]) + sm.bg[l][128 + col + e]);
        }
        put_all(l == 0 ? kSlotH0 : kSlotH1, rh);
        signal(1);
        if (l == 0 && do_sbias) {                                 // sbias[s] rode along with the gates of layer 0
          uint32_t v[8];
          tmem_ld_32x8(t_lane + kColSb + sub * 8, v);
          tmem_ld_wait8(v);
          if (ok) {
            float4* o = reinterpret_cast<float4*>(sbias + ((long long)s * B + b) * kDim + col);
            o[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
            o[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
          }
        }
        wait_acc();                                               // ---- E_c: h' = u*h + (1-u)*c
        uint32_t vc[8];
        tmem_ld_32x8(t_lane + kColC + sub * 8, vc);
        tmem_ld_wait8(vc);
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float c = tanh_fast(__uint_as_float(vc[e]) + sm.bc[l][col + e]);
          o[e] = fmaf(u[e], h[l][e] - c, c);                      // UNMASKED: the input of the layer above
          h[l][e] = m * o[e];                                     // state *= mask (model_hier.py:93)
        }
        if (l == 0) {
          put_all(kSlotH0, o);                                    // layer 1 reads [h0' | h1]; the h1 slot still holds h1
          signal(1);
        } else if (s + 1 < S) {
          put_x();                                                // next step: [x_{s+1} | m*h0' | m*h1']
          put_all(kSlotH0, h[0]);
          put_all(kSlotH1, h[1]);
          signal(2);
        }
      }
      // every accumulator wait has count 4 (all CTAs' MMAs of the phase): it doubles as "operand buffers free"
    }
    if (ok) {
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        float4* o = reinterpret_cast<float4*>(state_out + (long long)b * 256 + l * 128 + col);
        o[0] = make_float4(h[l][0], h[l][1], h[l][2], h[l][3]);
        o[1] = make_float4(h[l][4], h[l][5], h[l][6], h[l][7]);
      }
    }
  }
  tc_fence_before_sync();
  cluster_sync_all();                              // nobody exits while a peer may still write its smem / signal it
  if (warp == k3c::kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc<k3c::kTmemCols>(tmem);
  }
}

// fp32 TF-layout weights -> per-CTA blobs in the shared-memory layout [32 K chunks][n rows][8 bf16] per matrix
__global__ void k3_prepare_weights_cluster(const float* __restrict__ w_in_state, const float* const* __restrict__ w_dev /* gw0,cw0,gw1,cw1 */,
                                           __nv_bfloat16* __restrict__ out) {
  using namespace k3c;
  const int c = blockIdx.x / 5, mat = blockIdx.x % 5;            // CTA rank, matrix (sb, g0, c0, g1, c1)
  const int row_off = mat == 0 ? kOffSb : mat == 1 ? kOffG0 : mat == 2 ? kOffC0 : mat == 3 ? kOffG1 : kOffC1;
  const int n_rows = (mat == 1 || mat == 3) ? 64 : 32;
  const float* src = mat == 0 ? w_in_state : w_dev[mat - 1];
  const int ld = (mat == 1 || mat == 3) ? 256 : 128;
  __nv_bfloat16* dst = out + (size_t)c * (kBlobBytes / 2) + (size_t)row_off * 256;
  // one thread = one 16-byte K chunk of one row: 8 coalesced reads (a warp covers 32 consecutive output columns), one
  // 16-byte write; blockIdx.y picks 4 of the 32 K chunks
  const int i = threadIdx.x;
  if (i >= 4 * n_rows) return;
  const int kc = blockIdx.y * 4 + i / n_rows, n = i % n_rows;
  const int ocol = n_rows == 64 ? ((n >> 5) * 128 + c * kSlice + (n & 31)) : (c * kSlice + n);   // gates: r rows, then u rows
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = src ? src[(long long)(kc * 8 + e) * ld + ocol] : 0.f;
  *reinterpret_cast<uint4*>(dst + ((size_t)kc * n_rows + n) * 8) = pack8(v);
}

int32_t gru_sessions_bf16_cluster(const float* yp, const float* mask, const float* state_in, const float* const* gate_w,
                                  const float* const* gate_b, const float* const* cand_w, const float* const* cand_b,
                                  const float* w_in_state, int B, int S, float* state_pre, float* sbias, float* state_out,
                                  float* scratch, cudaStream_t st) {
  using namespace k3c;
  // scratch layout (same size as the streaming kernel's): [4 blobs of 112 KB][4 device pointers][768 bias floats]
  uint8_t* sc = reinterpret_cast<uint8_t*>(scratch);
  const size_t w_bytes = (size_t)kCl * kBlobBytes;
  static_assert(kCl * kBlobBytes == 14 * 128 * 128 * 2, "HTCN_GRU_SCRATCH_BYTES covers both kernels");
  const float** ptrs_dev = reinterpret_cast<const float**>(sc + w_bytes);
  float* bias_dev = reinterpret_cast<float*>(sc + w_bytes + 64);
  const float* ptrs[4] = {gate_w[0], cand_w[0], gate_w[1], cand_w[1]};
  HTCN_CUDA(cudaMemcpyAsync(ptrs_dev, ptrs, sizeof(ptrs), cudaMemcpyHostToDevice, st));
  for (int l = 0; l < 2; ++l) {
    HTCN_CUDA(cudaMemcpyAsync(bias_dev + l * 384, gate_b[l], 256 * 4, cudaMemcpyDeviceToDevice, st));
    HTCN_CUDA(cudaMemcpyAsync(bias_dev + l * 384 + 256, cand_b[l], 128 * 4, cudaMemcpyDeviceToDevice, st));
  }
  k3_prepare_weights_cluster<<<dim3(kCl * 5, 8), 256, 0, st>>>(w_in_state, ptrs_dev, reinterpret_cast<__nv_bfloat16*>(sc));
  HTCN_LAUNCH_CHECK("k3_prepare_weights_cluster");
  const size_t smem = sizeof(Smem) + 1024;
  // default (1): plain DSMEM stores published by one cluster-scope fence per warp and phase (146 us at B=4096, S=10);
  // HTCN_K3_CLUSTER=2: st.async with transaction bytes (154 us)
  const char* env = getenv("HTCN_K3_CLUSTER");
  auto kern = (env && atoi(env) == 2) ? k3_gru_bf16_cluster<true> : k3_gru_bf16_cluster<false>;
  HTCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<kCl * ceil_div(B, kM), kThreads, smem, st>>>(sc, yp, mask, state_in, bias_dev, B, S, sbias != nullptr, state_pre, sbias,
                                                     state_out);
  HTCN_LAUNCH_CHECK("k3_gru_bf16_cluster");
  return HTCN_OK;
}

}  // namespace htcn
