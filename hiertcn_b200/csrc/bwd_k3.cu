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
// The recurrence is a host loop of small launches (two skinny GEMMs + two elementwise kernels per cell call); all the
// weight-gradient products are deferred to the end, where they are GEMMs with S*B rows.
#include "train.cuh"

namespace htcn {

namespace {

struct CellPtrs {
  const float* r; const float* u; const float* c;   // [B,128] saved gates of this cell call
  const float* h;                                   // state before the call, row stride 256
  const float* mask;                                // [B] reset mask of this step
  const float* d_after;                             // dL/d(state after this step), row stride 256, or NULL (last step)
  const float* dx_upper;                            // dL/d(output) from the layer above, row stride 256, or NULL (top)
  float* dc_pre;                                    // [B,128]
  float* dg_pre;                                    // [B,256]: [drpre | dupre]
  float* carry;                                     // [B,128]: dh' u
};

__global__ void gru_bwd_a(int B, CellPtrs p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * kDim) return;
  const int b = (int)(i >> 7), c = (int)(i & 127);
  float dhn = p.d_after ? p.d_after[(long long)b * 256 + c] * p.mask[b] : 0.f;
  if (p.dx_upper) dhn += p.dx_upper[(long long)b * 256 + c];
  const float u = p.u[i], cc = p.c[i], h = p.h[(long long)b * 256 + c];
  const float du = dhn * (h - cc), dc = dhn * (1.f - u);
  p.dc_pre[i] = dc * (1.f - cc * cc);
  p.dg_pre[(long long)b * 256 + 128 + c] = du * u * (1.f - u);
  p.carry[i] = dhn * u;
}

// after [dx_c | drh] = dcpre Wc^T (tmpc): drpre, and start the output row [dx | dh] of this cell call
__global__ void gru_bwd_b(int B, const float* __restrict__ tmpc, const float* __restrict__ r, const float* __restrict__ h,
                          const float* __restrict__ carry, const float* __restrict__ sb_contrib, float* __restrict__ dg_pre,
                          float* __restrict__ out /* [B,256] */) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * kDim) return;
  const int b = (int)(i >> 7), c = (int)(i & 127);
  const float drh = tmpc[(long long)b * 256 + 128 + c];
  const float rr = r[i], hh = h[(long long)b * 256 + c];
  dg_pre[(long long)b * 256 + c] = drh * hh * rr * (1.f - rr);
  out[(long long)b * 256 + c] = tmpc[(long long)b * 256 + c];
  out[(long long)b * 256 + 128 + c] = carry[i] + drh * rr + sb_contrib[i];
}

// rh = r h and the new (pre-mask) state hn = u h + (1-u) c of every cell call, for the deferred weight gradients
__global__ void gru_bwd_setup(long long n /* S*B*128 */, const float* __restrict__ r, const float* __restrict__ u,
                              const float* __restrict__ c, const float* __restrict__ h /* stride 256 */,
                              float* __restrict__ rh, float* __restrict__ hn) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long row = i >> 7;
  const int col = (int)(i & 127);
  const float hh = h[row * 256 + col];
  rh[i] = r[i] * hh;
  if (hn) hn[i] = u[i] * hh + (1.f - u[i]) * c[i];
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
  float* tmpc = p; p += (long long)B * 256;
  float* carry = p;
  auto gate = [&](int s, int g, int which) { return gates_save + ((long long)(s * G + g) * 3 + which) * BD; };
  const int GH = G * kDim;
  HTCN_REQUIRE(GH == 256, "gru_backward: built for G*H == 256 (row stride of the saved states), got %d", GH);
  int32_t rc;

  // ---- setup: r*h, new states, and the sbias path's contribution to dL/dstate_pre for every step ---------------------
  for (int g = 0; g < G; ++g) {
    // gates of layer g are not contiguous over s (layout [S][G][3][B][128]) -> one launch per step
    for (int s = 0; s < S; ++s) {
      gru_bwd_setup<<<ceil_div(BD, 256), 256, 0, st>>>(BD, gate(s, g, 0), gate(s, g, 1), gate(s, g, 2),
                                                       state_pre + (long long)s * B * GH + g * kDim,
                                                       RH[g] + s * BD, (g + 1 < G) ? HN[g] + s * BD : nullptr);
      HTCN_LAUNCH_CHECK("gru_bwd_setup");
    }
    // SBC_g[s*B+b, :] = dsbias[s,b,:] @ W_in_state[g*128:(g+1)*128, :]^T
    rc = sgemm(true, SB, kDim, kDim, d_sbias, kDim, w_in_state + (long long)g * kDim * kDim, kDim, SBC[g], kDim, false, st);
    if (rc) return rc;
  }

  // ---- the recurrence, last step first ------------------------------------------------------------------------------
  for (int s = S - 1; s >= 0; --s) {
    for (int g = G - 1; g >= 0; --g) {
      CellPtrs c{};
      c.r = gate(s, g, 0); c.u = gate(s, g, 1); c.c = gate(s, g, 2);
      c.h = state_pre + (long long)s * B * GH + g * kDim;
      c.mask = mask + (long long)s * B;
      c.d_after = (s + 1 < S) ? OUT[g] + (long long)(s + 1) * B * 256 + 128 : nullptr;
      c.dx_upper = (g + 1 < G) ? OUT[g + 1] + (long long)s * B * 256 : nullptr;
      c.dc_pre = DC[g] + s * BD;
      c.dg_pre = DG[g] + (long long)s * B * 256;
      c.carry = carry;
      gru_bwd_a<<<ceil_div(BD, 256), 256, 0, st>>>(B, c);
      HTCN_LAUNCH_CHECK("gru_bwd_a");
      // [dx_c | drh] = dcpre [B,128] @ Wc^T  (Wc is [256,128])
      rc = sgemm(true, B, 256, kDim, c.dc_pre, kDim, cand_w_host[g], kDim, tmpc, 256, false, st);
      if (rc) return rc;
      float* out = OUT[g] + (long long)s * B * 256;
      gru_bwd_b<<<ceil_div(BD, 256), 256, 0, st>>>(B, tmpc, c.r, c.h, carry, SBC[g] + s * BD, c.dg_pre, out);
      HTCN_LAUNCH_CHECK("gru_bwd_b");
      // [dx | dh] += [drpre | dupre] [B,256] @ Wg^T  (Wg is [256,256])
      rc = sgemm(true, B, 256, 256, c.dg_pre, 256, gate_w_host[g], 256, out, 256, true, st);
      if (rc) return rc;
    }
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
  // dL/dyp[s] = dx of layer 0
  HTCN_CUDA(cudaMemcpy2DAsync(d_yp, sizeof(float) * kDim, OUT[0], sizeof(float) * 256, sizeof(float) * kDim, (size_t)SB,
                              cudaMemcpyDeviceToDevice, st));
  return HTCN_OK;
}
