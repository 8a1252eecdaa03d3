// K3 backward (fp32): back-propagation through the S session steps of the stacked GRU (customed_gru_cell.py:309-337,
// 1050-1073 as unrolled by model_hier.py:30-37,91,93), truncated at the batch boundary because the carried state is
// fed through a placeholder (model.py:44, run_hier_xing.py:291).
//
// For one cell call with input x, previous state h and saved r, u, c (h' = u h + (1-u) c), given dL/dh':
//     du = dh' (h - c)        dc = dh' (1 - u)          dcpre = dc (1 - c^2)        dupre = du u (1 - u)
//     [dx_c | drh] = dcpre Wc^T                         dr = drh h                  drpre = dr r (1 - r)
//     [dx_g | dh_g] = [drpre | dupre] Wg^T
//     dx = dx_c + dx_g        dh = dh' u + drh r + dh_g
//     dWc += [x | r h]^T dcpre,  dbc += sum dcpre,  dWg += [x | h]^T [drpre | dupre],  dbg += sum [drpre | dupre]
// The state before step s also feeds the TCN of slot s through sbias[s] = state_pre[s] W_in[128:] (model_hier.py:54-55),
// which adds dsbias[s] W_in[128:]^T to dL/dstate_pre[s] and state_pre^T dsbias to the gradient of W_in[128:].
//
// The recurrence is one persistent kernel (gru_bptt_kernel: a CTA per 32 users walks all S x G cell calls); all the
// weight-gradient products are deferred to the end, where they are GEMMs with S*B rows.
#include <cstdlib>

#include "train.cuh"

namespace htcn {

namespace {

// ---- the recurrence as ONE persistent kernel -----------------------------------------------------------------------
// Users are independent, so a CTA owns kBpMB = 32 of them and walks all S x G cell calls (last step first, top layer first)
// with the running gradients [dx | dh] of both layers in shared memory; the two skinny products of a cell call
// ([32 x 128] Wc^T and [32 x 256] Wg^T) are FFMA register tiles fed by double-buffered 16-deep weight slices out of L2.
// Replaces 4 launches per cell call (80 per step at S = 10, G = 2).  What the deferred weight-gradient GEMMs need (dcpre,
// [drpre | dupre]) is written out as before.
// kBpMB = users per CTA: 32 at large batches (16 with two CTAs per SM measured slower at 4096 users: 1.31 vs 1.21 ms), 8 when
// the batch is small (512 users per GPU in 8-way data-parallel training: 16 CTAs would leave 132 SMs idle)
constexpr int kBpThreads = 256;
constexpr int kBpKT = 16;            // contraction slice
template <int kBpMB>
struct BpttSmem {
  float outs[2][kBpMB][256];         // [layer][user][dx | dh] of the step being processed / the one after it
  float tmp[kBpMB][256];             // [dx_c | drh]
  float dg[kBpMB][256];              // [drpre | dupre]
  float dc[kBpMB][128];              // dcpre
  float carry[kBpMB][128];           // dh' u
  float wt[2][kBpKT][256];           // weight slices, transposed: wt[kk][n] = W[n][k0 + kk]
};
struct BpttArgs {
  const float* gates;                // [S][G][3][B][128]
  const float* state_pre;            // [S][B][256]
  const float* mask;                 // [S][B]
  const float* sbc[2];               // [S*B,128] per layer
  const float* wc[2];                // [256,128]
  const float* wg[2];                // [256,256]
  float* DC[2];                      // [S*B,128]
  float* DG[2];                      // [S*B,256]
  float* d_yp;                       // [S*B,128]
  int B, S;
};

// acc[i][j] += sum_k A[b0 + i][k] W[n4 + j][k],  k < K;  A in shared memory (row stride lda), W [256][K] in global memory
template <int K, int kBpTU, class Smem>
__device__ __forceinline__ void bp_gemm(const float* A, int lda, const float* __restrict__ W, float (&acc)[kBpTU][4], Smem& sm,
                                        int tid, int b0, int n4) {
  float4 pre[4];
  const float4* wrow = reinterpret_cast<const float4*>(W + (long long)tid * K);
#pragma unroll
  for (int i = 0; i < 4; ++i) pre[i] = __ldg(wrow + i);
  for (int k0 = 0; k0 < K; k0 += kBpKT) {
    float (*wt)[256] = sm.wt[(k0 / kBpKT) & 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      wt[4 * i + 0][tid] = pre[i].x; wt[4 * i + 1][tid] = pre[i].y; wt[4 * i + 2][tid] = pre[i].z; wt[4 * i + 3][tid] = pre[i].w;
    }
    __syncthreads();
    if (k0 + kBpKT < K) {
#pragma unroll
      for (int i = 0; i < 4; ++i) pre[i] = __ldg(wrow + (k0 + kBpKT) / 4 + i);
    }
#pragma unroll
    for (int kk = 0; kk < kBpKT; kk += 4) {
      float4 a[kBpTU];
#pragma unroll
      for (int i = 0; i < kBpTU; ++i) a[i] = *reinterpret_cast<const float4*>(A + (b0 + i) * lda + k0 + kk);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(&wt[kk + j][n4]);
#pragma unroll
        for (int i = 0; i < kBpTU; ++i) {
          const float av = j == 0 ? a[i].x : j == 1 ? a[i].y : j == 2 ? a[i].z : a[i].w;
          acc[i][0] = fmaf(av, w.x, acc[i][0]); acc[i][1] = fmaf(av, w.y, acc[i][1]);
          acc[i][2] = fmaf(av, w.z, acc[i][2]); acc[i][3] = fmaf(av, w.w, acc[i][3]);
        }
      }
    }
  }
}

template <int kBpMB>
__global__ void __launch_bounds__(kBpThreads, 1) gru_bptt_kernel(BpttArgs a) {
  constexpr int kBpTU = kBpMB / 4;     // users per thread tile
  extern __shared__ __align__(16) uint8_t bp_smem[];
  BpttSmem<kBpMB>& sm = *reinterpret_cast<BpttSmem<kBpMB>*>(bp_smem);
  const int tid = threadIdx.x;
  const int u0 = blockIdx.x * kBpMB;
  const int n4 = 4 * (tid & 63), b0 = kBpTU * (tid >> 6);
  const long long BD = (long long)a.B * kDim;
  constexpr int G = 2;
  for (int s = a.S - 1; s >= 0; --s) {
    for (int g = G - 1; g >= 0; --g) {
      const float* gr = a.gates + ((long long)(s * G + g) * 3 + 0) * BD;
      const float* gu = gr + BD;
      const float* gc = gu + BD;
      // ---- A: gate gradients of this cell call (float4 per thread and iteration, loads of all iterations in flight)
#pragma unroll
      for (int it4 = 0; it4 < (kBpMB * 32 + kBpThreads - 1) / kBpThreads; ++it4) {
        const int idx = tid + it4 * kBpThreads;
        if (kBpMB * 32 < kBpThreads && idx >= kBpMB * 32) break;
        const int b = idx >> 5, c = (idx & 31) * 4, ub = u0 + b;
        float4 dcp = make_float4(0.f, 0.f, 0.f, 0.f), dup = dcp, cr = dcp;
        if (ub < a.B) {
          const long long gi = (long long)ub * kDim + c, row = (long long)s * a.B + ub;
          const float4 u = *reinterpret_cast<const float4*>(gu + gi), cc = *reinterpret_cast<const float4*>(gc + gi);
          const float4 h = *reinterpret_cast<const float4*>(a.state_pre + row * 256 + g * kDim + c);
          float4 dhn = make_float4(0.f, 0.f, 0.f, 0.f);
          if (s + 1 < a.S) {
            const float mk = a.mask[(long long)s * a.B + ub];
            const float4 o = *reinterpret_cast<const float4*>(&sm.outs[g][b][128 + c]);
            dhn = make_float4(o.x * mk, o.y * mk, o.z * mk, o.w * mk);
          }
          if (g + 1 < G) {
            const float4 o = *reinterpret_cast<const float4*>(&sm.outs[g + 1][b][c]);
            dhn.x += o.x; dhn.y += o.y; dhn.z += o.z; dhn.w += o.w;
          }
#define HTCN_BP_A(f)                                        \
          {                                                 \
            const float du = dhn.f * (h.f - cc.f), dc = dhn.f * (1.f - u.f); \
            dcp.f = dc * (1.f - cc.f * cc.f);               \
            dup.f = du * u.f * (1.f - u.f);                 \
            cr.f = dhn.f * u.f;                             \
          }
          HTCN_BP_A(x) HTCN_BP_A(y) HTCN_BP_A(z) HTCN_BP_A(w)
#undef HTCN_BP_A
          *reinterpret_cast<float4*>(a.DC[g] + row * kDim + c) = dcp;
          *reinterpret_cast<float4*>(a.DG[g] + row * 256 + 128 + c) = dup;
        }
        *reinterpret_cast<float4*>(&sm.dc[b][c]) = dcp;
        *reinterpret_cast<float4*>(&sm.dg[b][128 + c]) = dup;
        *reinterpret_cast<float4*>(&sm.carry[b][c]) = cr;
      }
      __syncthreads();
      // ---- [dx_c | drh] = dcpre Wc^T
      float acc[kBpTU][4];
#pragma unroll
      for (int i = 0; i < kBpTU; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      bp_gemm<128, kBpTU>(&sm.dc[0][0], 128, a.wc[g], acc, sm, tid, b0, n4);
#pragma unroll
      for (int i = 0; i < kBpTU; ++i) *reinterpret_cast<float4*>(&sm.tmp[b0 + i][n4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      __syncthreads();
      // ---- B: drpre, and the start of this cell call's [dx | dh]
#pragma unroll
      for (int it4 = 0; it4 < (kBpMB * 32 + kBpThreads - 1) / kBpThreads; ++it4) {
        const int idx = tid + it4 * kBpThreads;
        if (kBpMB * 32 < kBpThreads && idx >= kBpMB * 32) break;
        const int b = idx >> 5, c = (idx & 31) * 4, ub = u0 + b;
        float4 drp = make_float4(0.f, 0.f, 0.f, 0.f), dh = drp;
        if (ub < a.B) {
          const long long gi = (long long)ub * kDim + c, row = (long long)s * a.B + ub;
          const float4 rr = *reinterpret_cast<const float4*>(gr + gi);
          const float4 hh = *reinterpret_cast<const float4*>(a.state_pre + row * 256 + g * kDim + c);
          const float4 sb = *reinterpret_cast<const float4*>(a.sbc[g] + row * kDim + c);
          const float4 drh = *reinterpret_cast<const float4*>(&sm.tmp[b][128 + c]);
          const float4 cr = *reinterpret_cast<const float4*>(&sm.carry[b][c]);
          drp = make_float4(drh.x * hh.x * rr.x * (1.f - rr.x), drh.y * hh.y * rr.y * (1.f - rr.y),
                            drh.z * hh.z * rr.z * (1.f - rr.z), drh.w * hh.w * rr.w * (1.f - rr.w));
          dh = make_float4(cr.x + drh.x * rr.x + sb.x, cr.y + drh.y * rr.y + sb.y, cr.z + drh.z * rr.z + sb.z,
                           cr.w + drh.w * rr.w + sb.w);
          *reinterpret_cast<float4*>(a.DG[g] + row * 256 + c) = drp;
        }
        *reinterpret_cast<float4*>(&sm.dg[b][c]) = drp;
        *reinterpret_cast<float4*>(&sm.outs[g][b][c]) = *reinterpret_cast<const float4*>(&sm.tmp[b][c]);
        *reinterpret_cast<float4*>(&sm.outs[g][b][128 + c]) = dh;
      }
      __syncthreads();
      // ---- [dx | dh] += [drpre | dupre] Wg^T
#pragma unroll
      for (int i = 0; i < kBpTU; ++i) {
        const float4 o = *reinterpret_cast<const float4*>(&sm.outs[g][b0 + i][n4]);
        acc[i][0] = o.x; acc[i][1] = o.y; acc[i][2] = o.z; acc[i][3] = o.w;
      }
      bp_gemm<256, kBpTU>(&sm.dg[0][0], 256, a.wg[g], acc, sm, tid, b0, n4);
#pragma unroll
      for (int i = 0; i < kBpTU; ++i) *reinterpret_cast<float4*>(&sm.outs[g][b0 + i][n4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      __syncthreads();
      if (g == 0) {                                              // dL/dyp[s] = dx of layer 0
#pragma unroll
        for (int it4 = 0; it4 < (kBpMB * 32 + kBpThreads - 1) / kBpThreads; ++it4) {
          const int idx = tid + it4 * kBpThreads;
          if (kBpMB * 32 < kBpThreads && idx >= kBpMB * 32) break;
          const int b = idx >> 5, c = (idx & 31) * 4, ub = u0 + b;
          if (ub < a.B)
            *reinterpret_cast<float4*>(a.d_yp + ((long long)s * a.B + ub) * kDim + c) = *reinterpret_cast<const float4*>(&sm.outs[0][b][c]);
        }
      }
    }
  }
}

// rh and hn of every cell call of every layer in one launch (gates [S][G][3][B][128])
__global__ void gru_bwd_setup_all(int S, int G, int B, const float* __restrict__ gates, const float* __restrict__ state_pre,
                                  float* rh0, float* rh1, float* hn0) {
  const long long BD = (long long)B * kDim;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)S * G * BD) return;
  const long long e = i % BD;
  const int sg = (int)(i / BD), s = sg / G, g = sg % G;
  const float* gp = gates + (long long)sg * 3 * BD;
  const float hh = state_pre[((long long)s * B + e / kDim) * 256 + g * kDim + e % kDim];
  const float r = gp[e], u = gp[BD + e], c = gp[2 * BD + e];
  (g == 0 ? rh0 : rh1)[s * BD + e] = r * hh;
  if (g == 0 && hn0) hn0[s * BD + e] = u * hh + (1.f - u) * c;
}

}  // namespace
}  // namespace htcn

// floats of scratch htcn_gru_backward needs
extern "C" int64_t htcn_gru_backward_scratch_floats(int32_t B, int32_t S, int32_t num_layer) {
  const int64_t SB = (int64_t)S * B;
  return (int64_t)num_layer * SB * 1024 + (int64_t)B * 384;
}

extern "C" int32_t htcn_gru_backward(const float* yp, const float* mask, const float* state_pre, const float* gates_save,
                                     const float* const* gate_w_host, const float* const* cand_w_host,
                                     int32_t num_layer, const float* w_in_state, int32_t B, int32_t S,
                                     const float* d_sbias, float* scratch, float* const* d_gate_w_host,
                                     float* const* d_gate_b_host, float* const* d_cand_w_host,
                                     float* const* d_cand_b_host, float* d_w_in_state, float* d_yp, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(yp && mask && state_pre && gates_save && gate_w_host && cand_w_host && w_in_state && d_sbias && scratch &&
                   d_gate_w_host && d_gate_b_host && d_cand_w_host && d_cand_b_host && d_w_in_state && d_yp,
               "gru_backward: NULL pointer");
  HTCN_REQUIRE(num_layer >= 1 && num_layer <= HTCN_MAX_GRU_LAYERS && B > 0 && S > 0, "gru_backward: G=%d B=%d S=%d",
               num_layer, B, S);
  cudaStream_t st = as_stream(stream);
  const int G = num_layer;
  const long long SB = (long long)S * B;
  const long long BD = (long long)B * kDim;
  // scratch carve-up (per layer: OUT [SB,256], DG [SB,256], DC, RH, SBC, HN [SB,128])
  float* OUT[HTCN_MAX_GRU_LAYERS]; float* DG[HTCN_MAX_GRU_LAYERS]; float* DC[HTCN_MAX_GRU_LAYERS];
  float* RH[HTCN_MAX_GRU_LAYERS]; float* SBC[HTCN_MAX_GRU_LAYERS]; float* HN[HTCN_MAX_GRU_LAYERS];
  float* p = scratch;
  for (int g = 0; g < G; ++g) {
    OUT[g] = p; p += SB * 256;
    DG[g] = p; p += SB * 256;
    DC[g] = p; p += SB * 128;
    RH[g] = p; p += SB * 128;
    SBC[g] = p; p += SB * 128;
    HN[g] = p; p += SB * 128;
  }
  const int GH = G * kDim;
  HTCN_REQUIRE(GH == 256, "gru_backward: built for G*H == 256 (row stride of the saved states), got %d", GH);
  int32_t rc;

  // ---- setup: r*h, new states, and the sbias path's contribution to dL/dstate_pre for every step ---------------------
  gru_bwd_setup_all<<<ceil_div((long long)S * G * BD, 256), 256, 0, st>>>(S, G, B, gates_save, state_pre, RH[0], RH[1], HN[0]);
  HTCN_LAUNCH_CHECK("gru_bwd_setup_all");
  for (int g = 0; g < G; ++g) {
    // SBC_g[s*B+b, :] = dsbias[s,b,:] @ W_in_state[g*128:(g+1)*128, :]^T
    rc = sgemm(true, SB, kDim, kDim, d_sbias, kDim, w_in_state + (long long)g * kDim * kDim, kDim, SBC[g], kDim, false, st);
    if (rc) return rc;
  }

  // ---- the recurrence, last step first: one persistent kernel, a CTA per 32 users ------------------------------------
  {
    BpttArgs ba{};
    ba.gates = gates_save; ba.state_pre = state_pre; ba.mask = mask; ba.d_yp = d_yp; ba.B = B; ba.S = S;
    for (int g = 0; g < G; ++g) {
      ba.sbc[g] = SBC[g]; ba.wc[g] = cand_w_host[g]; ba.wg[g] = gate_w_host[g]; ba.DC[g] = DC[g]; ba.DG[g] = DG[g];
    }
    const char* mbe = getenv("HTCN_BPTT_MB");                   // users per CTA: 8 or 32, default by batch size
    const int mb = mbe ? atoi(mbe) : (B <= 1024 ? 8 : 32);
    if (mb == 4) {
      HTCN_CUDA(cudaFuncSetAttribute(gru_bptt_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BpttSmem<4>)));
      gru_bptt_kernel<4><<<ceil_div(B, 4), kBpThreads, sizeof(BpttSmem<4>), st>>>(ba);
    } else if (mb == 8) {
      HTCN_CUDA(cudaFuncSetAttribute(gru_bptt_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BpttSmem<8>)));
      gru_bptt_kernel<8><<<ceil_div(B, 8), kBpThreads, sizeof(BpttSmem<8>), st>>>(ba);
    } else {
      HTCN_CUDA(cudaFuncSetAttribute(gru_bptt_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BpttSmem<32>)));
      gru_bptt_kernel<32><<<ceil_div(B, 32), kBpThreads, sizeof(BpttSmem<32>), st>>>(ba);
    }
    HTCN_LAUNCH_CHECK("gru_bptt_kernel");
  }

  // ---- deferred weight gradients: products over all S*B cell calls of a layer ---------------------------------------
  for (int g = 0; g < G; ++g) {
    const float* xin = (g == 0) ? yp : HN[g - 1];                 // [SB,128]
    const float* hprev = state_pre + g * kDim;                    // [SB,128] view with row stride 256
    for (int cb = 0; cb < 2; ++cb) {                               // column blocks of the gate weights
      rc = sgemm_tn_atomic(SB, xin, kDim, DG[g] + cb * kDim, 256, d_gate_w_host[g] + cb * kDim, 256, 0, 0, nullptr, st);
      if (rc) return rc;
      rc = sgemm_tn_atomic(SB, hprev, 256, DG[g] + cb * kDim, 256, d_gate_w_host[g] + (long long)kDim * 256 + cb * kDim,
                           256, 0, 0, nullptr, st);
      if (rc) return rc;
    }
    rc = sgemm_tn_atomic(SB, xin, kDim, DC[g], kDim, d_cand_w_host[g], kDim, 0, 0, nullptr, st);
    if (rc) return rc;
    rc = sgemm_tn_atomic(SB, RH[g], kDim, DC[g], kDim, d_cand_w_host[g] + (long long)kDim * kDim, kDim, 0, 0, nullptr, st);
    if (rc) return rc;
    rc = colsum_atomic(SB, DG[g], 256, 256, d_gate_b_host[g], st);
    if (rc) return rc;
    rc = colsum_atomic(SB, DC[g], kDim, kDim, d_cand_b_host[g], st);
    if (rc) return rc;
    // gradient of W_in[128 + g*128 : 128 + (g+1)*128, :] = state_pre[:, g]^T dsbias
    rc = sgemm_tn_atomic(SB, hprev, 256, d_sbias, kDim, d_w_in_state + (long long)g * kDim * kDim, kDim, 0, 0, nullptr, st);
    if (rc) return rc;
  }
  return HTCN_OK;                                                // (dL/dyp was written by the recurrence kernel)
}
