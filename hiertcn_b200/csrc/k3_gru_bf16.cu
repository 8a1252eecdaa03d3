// K3 (bf16 tier): GRU over sessions on the tensor cores.  One CTA owns 128 users (MMA M = 128) and runs all S
// steps; the recurrence never leaves the SM pair of buffers it needs:
//   * the GEMM operand [in | hidden] (128 x 256 bf16, no-swizzle K-major, 64 KB) lives in shared memory and is
//     rewritten in place by the epilogue warps between the five GEMM phases of a step,
//   * the fp32 recurrent state lives in TENSOR MEMORY: columns 256..383 = h0, 384..511 = h1 of the 128 users
//     (TMEM lane = user), read with tcgen05.ld and updated with tcgen05.st -- nothing of the recurrence touches
//     HBM/L2 except the step's input Yp[s] and the emitted sbias / state_pre rows,
//   * accumulators share TMEM columns 0..255: gates -> 0..255 (r | u); the candidate and the sbias product reuse
//     0..127 once r has been consumed, so the update gate's pre-activation is still there when h' is formed.
// Weights (bf16, [n][k] tiles of 32 KB, 14 per step, L2-resident) stream through a 3-stage TMA ring -- 192 KB per
// layer do not fit next to the operand, see DESIGN.md for the cluster variant that would make them resident.
//
// Step s (customed_gru_cell.py:309-337 per layer, :1050-1073 stacking; model_hier.py:54-55,91,93):
//   P_sb : sbias[s] = [h0 | h1] @ W_in[D:]                         (state BEFORE the session)
//   P_g0 : [r|u] = sigmoid([x | h0] Wg0 + bg0)          E: operand hidden half <- r * h0
//   P_c0 : c = tanh([x | r*h0] Wc0 + bc0)               E: h0' = u*h0 + (1-u)*c ; state <- m*h0' ; operand <- [h0' | h1]
//   P_g1 / P_c1 : the same for layer 1 with input h0'   E: state <- m*h1' ; operand <- [m*h0' | m*h1']
#include <cstdlib>

#include "common.cuh"
#include "sm100.cuh"

#ifndef HTCN_K3_DEFAULT_MODE
#define HTCN_K3_DEFAULT_MODE 5
#endif

namespace htcn {
using namespace sm100;

constexpr int kGM = 128;                          // users per CTA
constexpr int kGActBytes = 32 * kGM * 16;         // 32 x 16-byte K chunks (K = 256) x 128 rows = 64 KB
constexpr int kGTileBytes = 2 * 128 * 128;        // one weight tile [128 n][128 k] bf16 = 32 KB
constexpr int kGStages = 3;
constexpr int kGEpiWarps = 16;                    // 4 TMEM lane quarters x 4 column groups of 32: a thread owns 32 columns of one user
constexpr int kGThreads = 32 * (kGEpiWarps + 2);  // warps 0-15 epilogue, 16 = TMA producer, 17 = MMA issuer
constexpr int kGProducerWarp = kGEpiWarps, kGMmaWarp = kGEpiWarps + 1;
constexpr int kTilesSb = 2, kTilesGates = 4, kTilesCand = 2;

struct alignas(1024) K3Smem {
  uint8_t w[kGStages][kGTileBytes];               // 96 KB
  uint8_t act[kGActBytes];                        // 64 KB
  float bg[2][256];
  float bc[2][128];
  uint64_t w_full[kGStages], w_empty[kGStages], acc_ready, act_ready;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t make_desc_gru_act(uint32_t smem_addr) {   // rows 16 B apart, K chunks kGM*16 B apart
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((kGM * 16) >> 4) << 16;
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

// write 8 consecutive K elements (one 16-byte chunk) of this thread's row
__device__ __forceinline__ void put_chunk(uint8_t* act_row, int chunk, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(act_row + chunk * (kGM * 16)) =
      make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
// operand half (0 = input, 1 = hidden), columns cg*32..+31 <- bf16(src[..]) (src may be NULL -> zeros)
__device__ __forceinline__ void put_half(uint8_t* act_row, int half, const float* src, int cg) {
#pragma unroll
  for (int c = cg * 4; c < cg * 4 + 4; ++c) {
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (src) {
      const float4 a = *reinterpret_cast<const float4*>(src + c * 8), b = *reinterpret_cast<const float4*>(src + c * 8 + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    put_chunk(act_row, half * 16 + c, v);
  }
}

constexpr uint32_t kColH = 256;                   // TMEM columns of the fp32 state: [h0 128 | h1 128]

// this thread's 32 state values (layer l, columns cc*32..+31) <-> TMEM
__device__ __forceinline__ void ld_state(uint32_t t_lane, int l, int cc, float (&h)[32]) {
  uint32_t v[32];
  tmem_ld_32x32(t_lane + kColH + l * 128 + cc * 32, v);
  tmem_ld_wait(v);
#pragma unroll
  for (int i = 0; i < 32; ++i) h[i] = __uint_as_float(v[i]);
}
__device__ __forceinline__ void st_state(uint32_t t_lane, int l, int cc, const float (&h)[32]) {
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(h[i]);
  tmem_st_32x32(t_lane + kColH + l * 128 + cc * 32, v);
}
// operand half, columns cc*32..+31 <- bf16(state layer l)
__device__ __forceinline__ void put_half_state(uint8_t* act_row, int half, uint32_t t_lane, int l, int cc) {
  {
    float h[32];
    ld_state(t_lane, l, cc, h);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float v[8] = {h[q * 8], h[q * 8 + 1], h[q * 8 + 2], h[q * 8 + 3], h[q * 8 + 4], h[q * 8 + 5], h[q * 8 + 6], h[q * 8 + 7]};
      put_chunk(act_row, half * 16 + cc * 4 + q, v);
    }
  }
}

__global__ void __launch_bounds__(kGThreads, 1)
k3_gru_bf16(const __grid_constant__ CUtensorMap tmap_w, const float* __restrict__ yp, const float* __restrict__ mask,
            const float* __restrict__ state_in, const float* __restrict__ bias_all /* [bg0 256][bc0 128][bg1 256][bc1 128] */,
            int B, int S, int do_sbias, float* __restrict__ state_pre, float* __restrict__ sbias,
            float* __restrict__ state_out) {
  extern __shared__ uint8_t smem_raw[];
  auto& sm = *reinterpret_cast<K3Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_step = (do_sbias ? kTilesSb : 0) + 2 * (kTilesGates + kTilesCand);
  const int phases_per_step = (do_sbias ? 1 : 0) + 4;

  if (tid == 0) {
    prefetch_tmap(&tmap_w);
    for (int s = 0; s < kGStages; ++s) {
      mbar_init(&sm.w_full[s], 1);
      mbar_init(&sm.w_empty[s], 1);
    }
    mbar_init(&sm.acc_ready, 1);
    mbar_init(&sm.act_ready, 32 * kGEpiWarps);
    fence_barrier_init();
  }
  for (int i = tid; i < 768; i += kGThreads) {
    const int l = i / 384, j = i % 384;
    if (j < 256) sm.bg[l][j] = bias_all[i];
    else sm.bc[l][j - 256] = bias_all[i];
  }
  if (warp == kGMmaWarp) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, sm.tmem_base, 0);

  if (warp == kGProducerWarp) {
    // ===================== weight producer =====================
    if (lane == 0) {
      long long n = 0;
      for (int s = 0; s < S; ++s) {
        for (int j = do_sbias ? 0 : kTilesSb; j < 14; ++j, ++n) {      // global tile order: [sb 2][g0 4][c0 2][g1 4][c1 2]
          const int st = (int)(n % kGStages);
          mbar_wait_relaxed(&sm.w_empty[st], (uint32_t)(((n / kGStages) & 1) ^ 1));
          mbar_arrive_expect_tx(&sm.w_full[st], kGTileBytes);
          tma_load_2d(sm.w[st], &tmap_w, 0, j * 128, &sm.w_full[st]);
          tma_load_2d(sm.w[st] + kGTileBytes / 2, &tmap_w, 64, j * 128, &sm.w_full[st]);
        }
      }
      (void)tiles_per_step;
    }
  } else if (warp == kGMmaWarp) {
    // ===================== MMA issuer =====================
    // the whole warp runs the (warp-uniform) waits and descriptor arithmetic, one elected lane issues: under an
    // `if (lane == 0)` branch every tcgen05.mma is wrapped in an ELECT / R2UR.BROADCAST waterfall (~100 clk per MMA)
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(kGM, 128);
      const uint32_t act0 = smem_u32(sm.act);
      long long n = 0, n_act = 0;
      // one weight tile = (N half, K half): D columns d_col, A chunks k_half*16..+15
      auto tile_mma = [&](uint32_t d_col, int k_half, bool first) {
        const int st = (int)(n % kGStages);
        mbar_wait(&sm.w_full[st], (uint32_t)((n / kGStages) & 1));
        tc_fence_after_sync();
        const uint32_t w_base = smem_u32(sm.w[st]);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t da = make_desc_gru_act(act0 + (uint32_t)(k_half * 16 + 2 * k) * (kGM * 16));
          const uint64_t db = make_desc_k_sw128(w_base + (k >> 2) * (kGTileBytes / 2) + (k & 3) * 32);
          if (leader) umma_bf16(tmem + d_col, da, db, idesc, !(first && k == 0));
        }
        if (leader) umma_commit(&sm.w_empty[st]);
        ++n;
      };
      auto wait_operand = [&]() {
        mbar_wait(&sm.act_ready, (uint32_t)(n_act & 1));
        tc_fence_after_sync();
        ++n_act;
      };
      for (int s = 0; s < S; ++s) {
        if (do_sbias) {
          wait_operand();
          tile_mma(0, 0, true);
          tile_mma(0, 1, false);
          if (leader) umma_commit(&sm.acc_ready);
        }
        for (int l = 0; l < 2; ++l) {
          wait_operand();                                   // gates: N = 256 as two column halves
          tile_mma(0, 0, true);
          tile_mma(0, 1, false);
          tile_mma(128, 0, true);
          tile_mma(128, 1, false);
          if (leader) umma_commit(&sm.acc_ready);
          wait_operand();                                   // candidate: N = 128
          tile_mma(0, 0, true);                             // reuses the r columns, consumed by E_g
          tile_mma(0, 1, false);
          if (leader) umma_commit(&sm.acc_ready);
        }
      }
    }
  } else {
    // ===================== epilogue: thread = 32 columns (cc) of one user row =====================
    const int quarter = warp & 3, cc = warp >> 2;
    const int r = quarter * 32 + lane;
    const int b = blockIdx.x * kGM + r;
    const bool ok = b < B;
    uint8_t* act_row = sm.act + r * 16;
    const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16);
    long long n_acc = 0;
    auto operand_ready = [&]() {
      tc_fence_before_sync();
      fence_proxy_async_smem();
      mbar_arrive(&sm.act_ready);
    };
    auto wait_acc = [&]() {
      mbar_wait(&sm.acc_ready, (uint32_t)(n_acc & 1));
      tc_fence_after_sync();
      ++n_acc;
    };
    // prologue: TMEM state <- state_in (rows beyond B hold zeros)
    for (int l = 0; l < 2; ++l) {
        float h[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 a = ok ? __ldg(reinterpret_cast<const float4*>(state_in + (long long)b * 256 + l * 128 + cc * 32) + i)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
          h[i * 4] = a.x; h[i * 4 + 1] = a.y; h[i * 4 + 2] = a.z; h[i * 4 + 3] = a.w;
        }
        st_state(t_lane, l, cc, h);
      }
    tmem_st_wait();
    for (int s = 0; s < S; ++s) {
      const float m = ok ? __ldg(mask + (long long)s * B + b) : 0.f;
      const float* x = ok ? yp + ((long long)s * B + b) * kDim : nullptr;
      if (state_pre) {
        for (int l = 0; l < 2; ++l) {
            float h[32];
            ld_state(t_lane, l, cc, h);
            if (ok) {
              float4* o = reinterpret_cast<float4*>(state_pre + ((long long)s * B + b) * 256 + l * 128 + cc * 32);
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = make_float4(h[i * 4], h[i * 4 + 1], h[i * 4 + 2], h[i * 4 + 3]);
            }
          }
      }
      if (do_sbias) {
        put_half_state(act_row, 0, t_lane, 0, cc);
        put_half_state(act_row, 1, t_lane, 1, cc);
        operand_ready();
        wait_acc();                                          // ---- E_sb: sbias[s] out, operand <- [x | h0]
        {
          uint32_t v[32];
          tmem_ld_32x32(t_lane + cc * 32, v);
          tmem_ld_wait(v);
          if (ok) {
            float4* o = reinterpret_cast<float4*>(sbias + ((long long)s * B + b) * kDim + cc * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              o[q] = make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]), __uint_as_float(v[q * 4 + 2]),
                                 __uint_as_float(v[q * 4 + 3]));
          }
        }
      }
      put_half(act_row, 0, x, cc);
      put_half_state(act_row, 1, t_lane, 0, cc);
      operand_ready();
      for (int l = 0; l < 2; ++l) {
        wait_acc();                                            // ---- E_g: operand hidden half <- r * h
        {
          uint32_t v[32];
          float h[32];
          tmem_ld_32x32(t_lane + cc * 32, v);                  // r pre-activations, columns cc*32..+31
          tmem_ld_wait(v);
          ld_state(t_lane, l, cc, h);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int col = cc * 32 + q * 8 + e;
              o[e] = sigmoid_fast(__uint_as_float(v[q * 8 + e]) + sm.bg[l][col]) * h[q * 8 + e];
            }
            put_chunk(act_row, 16 + cc * 4 + q, o);
          }
        }
        operand_ready();
        wait_acc();                                            // ---- E_c: h' = u*h + (1-u)*c
        {
          uint32_t vc[32], vu[32];
          float h[32];
          tmem_ld_32x32(t_lane + cc * 32, vc);                 // candidate pre-activations (columns 0..127, reused)
          tmem_ld_32x32(t_lane + 128 + cc * 32, vu);           // update-gate pre-activations (still in TMEM)
          tmem_ld_wait(vc);
          tmem_ld_wait(vu);
          ld_state(t_lane, l, cc, h);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int col = cc * 32 + q * 8 + e;
              const float c = tanh_fast(__uint_as_float(vc[q * 8 + e]) + sm.bc[l][col]);
              const float u = sigmoid_fast(__uint_as_float(vu[q * 8 + e]) + sm.bg[l][128 + col]);
              o[e] = fmaf(u, h[q * 8 + e] - c, c);             // u*h + (1-u)*c, UNMASKED: the input of the layer above
              h[q * 8 + e] = m * o[e];                         // state *= mask (model_hier.py:93)
            }
            if (l == 0) put_chunk(act_row, cc * 4 + q, o);     // layer 1 input half
          }
          st_state(t_lane, l, cc, h);
        }
        tmem_st_wait();
        if (l == 0) {
          put_half_state(act_row, 1, t_lane, 1, cc);           // [h0' | h1]
          operand_ready();
        }
      }
      // after layer 1: the next phase (sbias of step s+1 or [x | h0] of step s+1) is staged at the top of the loop
    }
    for (int l = 0; l < 2; ++l) {
        float h[32];
        ld_state(t_lane, l, cc, h);
        if (ok) {
          float4* o = reinterpret_cast<float4*>(state_out + (long long)b * 256 + l * 128 + cc * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = make_float4(h[i * 4], h[i * 4 + 1], h[i * 4 + 2], h[i * 4 + 3]);
        }
      }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kGMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem);
  }
}

// fp32 TF-layout weights -> 14 bf16 [n][k] tiles: [sb k0,k1][per layer: gates (n0k0,n0k1,n1k0,n1k1), cand (k0,k1)]
__global__ void k3_prepare_weights(const float* __restrict__ w_in_state, const float* const* __restrict__ w_dev /* gw0,cw0,gw1,cw1 */,
                                   __nv_bfloat16* __restrict__ out) {
  const int j = blockIdx.x;
  const float* src;
  int ld, n0, k0;
  if (j < 2) { src = w_in_state; ld = 128; n0 = 0; k0 = j * 128; }
  else {
    const int l = (j - 2) / 6, t = (j - 2) % 6;
    if (t < 4) { src = w_dev[2 * l]; ld = 256; n0 = (t >> 1) * 128; k0 = (t & 1) * 128; }
    else { src = w_dev[2 * l + 1]; ld = 128; n0 = 0; k0 = (t - 4) * 128; }
  }
  for (int i = threadIdx.x; i < 128 * 128; i += blockDim.x) {
    const int n = i / 128, k = i % 128;
    out[(long long)j * 128 * 128 + i] = __float2bfloat16_rn(w_in_state || j >= 2 ? src[(long long)(k0 + k) * ld + n0 + n] : 0.f);
  }
}

int32_t gru_sessions_bf16_cluster(const float* yp, const float* mask, const float* state_in, const float* const* gate_w,
                                  const float* const* gate_b, const float* const* cand_w, const float* const* cand_b,
                                  const float* w_in_state, int B, int S, float* state_pre, float* sbias, float* state_out,
                                  float* scratch, cudaStream_t st);

int32_t gru_sessions_bf16_t(const float* yp, const float* mask, const float* state_in, const float* const* gate_w,
                            const float* const* gate_b, const float* const* cand_w, const float* const* cand_b,
                            const float* w_in_state, int B, int S, float* state_pre, float* sbias, float* state_out,
                            float* scratch, int variant, cudaStream_t st, float* gates_save = nullptr);

int32_t gru_sessions_bf16_w(const float* yp, const float* mask, const float* state_in, const float* const* gate_w,
                            const float* const* gate_b, const float* const* cand_w, const float* const* cand_b,
                            const float* w_in_state, int B, int S, float* state_pre, float* sbias, float* state_out,
                            float* scratch, int variant, cudaStream_t st);

// scratch layout of the kernels that stream [n][k] weight tiles: [14 bf16 weight tiles][4 device pointers][768 bias floats]
int32_t k3_prepare_stream_weights(const float* const* gate_w, const float* const* gate_b, const float* const* cand_w,
                                  const float* const* cand_b, const float* w_in_state, float* scratch, cudaStream_t st,
                                  __nv_bfloat16** w_out, float** bias_out) {
  uint8_t* sc = reinterpret_cast<uint8_t*>(scratch);
  __nv_bfloat16* w_bf16 = reinterpret_cast<__nv_bfloat16*>(sc);
  const size_t w_bytes = (size_t)14 * 128 * 128 * 2;
  const float** ptrs_dev = reinterpret_cast<const float**>(sc + w_bytes);
  float* bias_dev = reinterpret_cast<float*>(sc + w_bytes + 64);
  const float* ptrs[4] = {gate_w[0], cand_w[0], gate_w[1], cand_w[1]};
  HTCN_CUDA(cudaMemcpyAsync(ptrs_dev, ptrs, sizeof(ptrs), cudaMemcpyHostToDevice, st));
  for (int l = 0; l < 2; ++l) {
    HTCN_CUDA(cudaMemcpyAsync(bias_dev + l * 384, gate_b[l], 256 * 4, cudaMemcpyDeviceToDevice, st));
    HTCN_CUDA(cudaMemcpyAsync(bias_dev + l * 384 + 256, cand_b[l], 128 * 4, cudaMemcpyDeviceToDevice, st));
  }
  k3_prepare_weights<<<14, 256, 0, st>>>(w_in_state, ptrs_dev, w_bf16);
  HTCN_LAUNCH_CHECK("k3_prepare_weights");
  *w_out = w_bf16;
  *bias_out = bias_dev;
  return HTCN_OK;
}

// Weight preparation of the users-on-N kernels (k3_gru_t.cu, k3_gru_w.cu): ONE launch writes the 14 bf16 tiles and the bias
// block (the pointers travel as kernel arguments: no memcpy nodes -- the five cudaMemcpyAsync of k3_prepare_stream_weights cost
// more than the kernel they feed).   scratch layout: [14 bf16 tiles][768 bias floats]
struct K3RepArgs {
  const float* w[5];     // W_in[D:], gates 0, candidate 0, gates 1, candidate 1
  const float* b[4];     // gates 0, candidate 0, gates 1, candidate 1
};
__global__ void k3_prepare_weights_rep(K3RepArgs a, __nv_bfloat16* __restrict__ out, float* __restrict__ bias_out) {
  const int j = blockIdx.x;
  const float* src;
  int ld, n0, k0;
  if (j < 2) { src = a.w[0]; ld = 128; n0 = 0; k0 = j * 128; }
  else {
    const int l = (j - 2) / 6, t = (j - 2) % 6;
    if (t < 4) { src = a.w[1 + 2 * l]; ld = 256; n0 = (t >> 1) * 128; k0 = (t & 1) * 128; }
    else { src = a.w[2 + 2 * l]; ld = 128; n0 = 0; k0 = (t - 4) * 128; }
  }
  // written as the SHARED-MEMORY IMAGE of the two 16 KB sub-tiles (k halves of 64) of the tile: rows of 128 B with the 128-byte
  // swizzle already applied (16-byte chunk index ^= row % 8), so a kernel fetches a sub-tile with ONE 1-D bulk copy
  // (cp.async.bulk, no tensor map, no per-row address generation) and reads it with the SWIZZLE_128B UMMA descriptor
  __nv_bfloat16* dst = out + (long long)j * 128 * 128;
  // thread = (n, 8 consecutive k): coalesced reads along n, one 16-byte store
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < 128 * 16; i += blockDim.x * gridDim.y) {
    const int n = i & 127, kc = i >> 7;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = src ? __ldg(src + (long long)(k0 + kc * 8 + e) * ld + n0 + n) : 0.f;
    const int hf = kc >> 3, c = (kc & 7) ^ (n & 7);
    *reinterpret_cast<uint4*>(dst + hf * (128 * 64) + n * 64 + c * 8) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
  if (j == 0 && blockIdx.y == 0)
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
      const int l = i / 384, q = i % 384;
      bias_out[i] = q < 256 ? a.b[2 * l][q] : a.b[2 * l + 1][q - 256];
    }
}

int32_t k3_prepare_rep(const float* const* gate_w, const float* const* gate_b, const float* const* cand_w,
                       const float* const* cand_b, const float* w_in_state, float* scratch, cudaStream_t st,
                       const uint8_t** w_out, float** bias_out) {
  K3RepArgs a;
  a.w[0] = w_in_state;
  for (int l = 0; l < 2; ++l) {
    a.w[1 + 2 * l] = gate_w[l];
    a.w[2 + 2 * l] = cand_w[l];
    a.b[2 * l] = gate_b[l];
    a.b[2 * l + 1] = cand_b[l];
  }
  __nv_bfloat16* w_bf16 = reinterpret_cast<__nv_bfloat16*>(scratch);
  float* bias_dev = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(scratch) + (size_t)14 * 128 * 128 * 2);
  k3_prepare_weights_rep<<<dim3(14, 8), 256, 0, st>>>(a, w_bf16, bias_dev);
  HTCN_LAUNCH_CHECK("k3_prepare_weights_rep");
  *w_out = reinterpret_cast<const uint8_t*>(w_bf16);
  *bias_out = bias_dev;
  return HTCN_OK;
}

int32_t gru_sessions_bf16(const float* yp, const float* mask, const float* state_in, const float* const* gate_w,
                          const float* const* gate_b, const float* const* cand_w, const float* const* cand_b,
                          const float* w_in_state, int B, int S, float* state_pre, float* sbias, float* state_out,
                          float* scratch, cudaStream_t st, float* gates_save) {
  // training forward (gates_save != NULL): the users-on-N kernel with the gate activations written out
  if (gates_save)
    return gru_sessions_bf16_t(yp, mask, state_in, gate_w, gate_b, cand_w, cand_b, w_in_state, B, S, state_pre, sbias,
                               state_out, scratch, 2, st, gates_save);
  // HTCN_K3_CLUSTER picks the kernel: 0 = this file's single-CTA kernel (128 users on the MMA M axis, weights streamed
  // from L2); 1 / 2 = the 4-CTA cluster kernel with shared-memory-resident weight slices (k3_gru_cluster.cu; plain DSMEM
  // stores or st.async); 3, 4 = the users-on-N kernel (k3_gru_t.cu: hidden units on the MMA M axis, 32 users per CTA on N,
  // no exchange between CTAs) in its two shared-memory budgets; 5 (default) = the same with the three weight tiles the
  // recurrence waits for held in TENSOR MEMORY as the MMA's A operand (68 us at B=4096, S=10 against 79); 6, 7 = the
  // wavefront form (k3_gru_w.cu: layer 0 of step t+1 beside layer 1 of step t)
  const char* cl = getenv("HTCN_K3_CLUSTER");
  const int mode = cl ? atoi(cl) : HTCN_K3_DEFAULT_MODE;
  if (mode >= 6)
    return gru_sessions_bf16_w(yp, mask, state_in, gate_w, gate_b, cand_w, cand_b, w_in_state, B, S, state_pre, sbias,
                               state_out, scratch, mode - 6, st);
  if (mode >= 3)
    return gru_sessions_bf16_t(yp, mask, state_in, gate_w, gate_b, cand_w, cand_b, w_in_state, B, S, state_pre, sbias,
                               state_out, scratch, mode - 3, st);
  if (mode != 0)
    return gru_sessions_bf16_cluster(yp, mask, state_in, gate_w, gate_b, cand_w, cand_b, w_in_state, B, S, state_pre, sbias,
                                     state_out, scratch, st);
  __nv_bfloat16* w_bf16;
  float* bias_dev;
  int32_t rc = k3_prepare_stream_weights(gate_w, gate_b, cand_w, cand_b, w_in_state, scratch, st, &w_bf16, &bias_dev);
  if (rc) return rc;
  CUtensorMap tw;
  rc = make_tmap_bf16(&tw, w_bf16, (uint64_t)14 * 128, kDim, kDim, 64, 128, 128);
  if (rc) return rc;
  const size_t smem = sizeof(K3Smem) + 1024;
  HTCN_CUDA(cudaFuncSetAttribute(k3_gru_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k3_gru_bf16<<<ceil_div(B, kGM), kGThreads, smem, st>>>(tw, yp, mask, state_in, bias_dev, B, S, sbias != nullptr, state_pre,
                                                        sbias, state_out);
  HTCN_LAUNCH_CHECK("k3_gru_bf16");
  return HTCN_OK;
}

}  // namespace htcn
