// K1 backward (row scatter-add into the dense embedding gradient), the per-row loss weights of the reference's two-level
// mean, the fused TF-flavoured Adam update, and the refresh of the bf16 scoring table from its fp32 master.
#include "train.cuh"

namespace htcn {
namespace {

// One warp per position (b,t).  x side: dE[x] += dXe[b,t]  (id 0 reads no row, model.py:59-61, so it gets no gradient).
// y side: Yp[s,b] = mean_t E[y[b,s,t]] + b_emb (model_hier.py:83-85)  =>  dE[y] += dYp[s,b] / n_{b,s}.
__global__ void gather_backward_kernel(const float* __restrict__ d_xe, const float* __restrict__ d_yp,
                                       const int* __restrict__ x_id, const int* __restrict__ y_id, SlotTable slots, int B,
                                       int T, int item_num, float* __restrict__ d_emb) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)B * T) return;
  const int b = (int)(w / T), t = (int)(w % T);
  const int x = x_id[w], y = y_id[w];
  if (x > 0 && x < item_num) {
    const float4 g = *reinterpret_cast<const float4*>(d_xe + w * kDim + lane * 4);
    atomicAdd(reinterpret_cast<float4*>(d_emb + (long long)x * kDim + lane * 4), g);
  }
  if (y > 0 && y < item_num) {
    int s = 0;
    while (s + 1 < slots.n && slots.off[s + 1] <= t) ++s;
    int n = 0;
    for (int tt = slots.off[s] + lane; tt < slots.off[s + 1]; tt += 32) n += (y_id[(long long)b * T + tt] > 0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    const float inv = 1.0f / (float)n;
    float4 g = *reinterpret_cast<const float4*>(d_yp + ((long long)s * B + b) * kDim + lane * 4);
    g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
    atomicAdd(reinterpret_cast<float4*>(d_emb + (long long)y * kDim + lane * 4), g);
  }
}

// g_row[row_of[b,t]] = 1 / (n_b + 1e-6), n_b = #{t: y[b,t] > 0}   (model.py:111-116; the 1/user_count of :117 is applied
// by the optimiser so that data-parallel ranks can sum their gradients first)
__global__ void loss_row_weights_kernel(const int* __restrict__ y_id, const int* __restrict__ row_of, int T,
                                        float* __restrict__ g_row) {
  const int b = blockIdx.x;
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  int n = 0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) n += (y_id[(long long)b * T + t] > 0);
  if (n) atomicAdd(&cnt, n);
  __syncthreads();
  const float w = 1.0f / ((float)cnt + 1e-6f);
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int q = row_of[(long long)b * T + t];
    if (q >= 0) g_row[q] = w;
  }
}

// tf.train.AdamOptimizer: epsilon outside the bias-corrected root; lr_t carries the bias corrections
__global__ void adam_kernel(long long n4, float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                            float4* __restrict__ v, float lr_t, float b1, float b2, float eps,
                            const float* __restrict__ grad_div, int zero_grad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  // a batch without a single scored user (user_count == 0) has no loss: the reference's 0/0 would turn every weight
  // and both moments into NaN for good -- skip the update (the step counter of the caller still advances)
  const float div = grad_div ? *grad_div : 1.0f;
  if (!(div > 0.f)) {
    if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float sc = 1.0f / div;
  float4 gg = g[i], mm = m[i], vv = v[i], pp = p[i];
#define HTCN_ADAM1(c)                                 \
  {                                                   \
    const float gr = gg.c * sc;                       \
    mm.c = b1 * mm.c + (1.f - b1) * gr;               \
    vv.c = b2 * vv.c + (1.f - b2) * gr * gr;          \
    pp.c -= lr_t * mm.c / (sqrtf(vv.c) + eps);        \
  }
  HTCN_ADAM1(x) HTCN_ADAM1(y) HTCN_ADAM1(z) HTCN_ADAM1(w)
#undef HTCN_ADAM1
  p[i] = pp; m[i] = mm; v[i] = vv;
  if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// W_out^T master [N,128] f32 (+ b_out) -> the bf16 scoring layout [N,144] = [128 weights | b_hi | b_lo | 0 ...]
__global__ void refresh_wout_bf16_kernel(const float* __restrict__ wt, const float* __restrict__ b, int N,
                                         __nv_bfloat16* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one output pair
  constexpr int pairs = kWtPitchBf16 / 2;                                     // 72
  if (i >= (long long)N * pairs) return;
  const long long n = i / pairs;
  const int c = (int)(i % pairs) * 2;
  float lo = 0.f, hi = 0.f;
  if (c < kDim) {
    const float2 w = *reinterpret_cast<const float2*>(wt + n * kDim + c);
    lo = w.x; hi = w.y;
  } else if (c == kDim || c == kDim + 2) {        // [b_hi, b_lo | b_hi, b_lo, 1, 1, 1 | 0...]: see prepare_wout_kernel
    const float bv = b[n];
    lo = __bfloat162float(__float2bfloat16_rn(bv));
    hi = bv - lo;
  } else if (c == kDim + 4) {
    lo = hi = 1.f;
  } else if (c == kDim + 6) {
    lo = 1.f;
  }
  *reinterpret_cast<uint32_t*>(out + n * kWtPitchBf16 + c) = pack_bf16x2(lo, hi);
}

}  // namespace
}  // namespace htcn

using namespace htcn;

extern "C" int32_t htcn_gather_backward(const float* d_xe, const float* d_yp, const int32_t* x_id, const int32_t* y_id,
                                        const int32_t* slot_off_host, int32_t B, int32_t T, int32_t S, int32_t item_num,
                                        float* d_emb, float* d_emb_bias, void* stream) {
  HTCN_REQUIRE(d_xe && d_yp && x_id && y_id && slot_off_host && d_emb && d_emb_bias, "gather_backward: NULL pointer");
  HTCN_REQUIRE(B > 0 && T > 0 && S > 0 && S <= HTCN_MAX_SLOTS && item_num > 0, "gather_backward: B=%d T=%d S=%d", B, T, S);
  SlotTable slots;
  slots.n = S;
  for (int i = 0; i <= S; ++i) slots.off[i] = slot_off_host[i];
  HTCN_REQUIRE(slots.off[0] == 0 && slots.off[S] == T, "gather_backward: slot_off does not span T");
  cudaStream_t st = as_stream(stream);
  const long long warps = (long long)B * T;
  gather_backward_kernel<<<ceil_div(warps * 32, 256), 256, 0, st>>>(d_xe, d_yp, x_id, y_id, slots, B, T, item_num, d_emb);
  HTCN_LAUNCH_CHECK("gather_backward_kernel");
  return colsum_atomic((long long)S * B, d_yp, kDim, kDim, d_emb_bias, st);      // b_emb enters every Yp[s,b] once
}

extern "C" int32_t htcn_loss_row_weights(const int32_t* y_id, const int32_t* row_of, int32_t B, int32_t T, float* g_row,
                                         void* stream) {
  HTCN_REQUIRE(y_id && row_of && g_row && B > 0 && T > 0, "loss_row_weights: bad args");
  loss_row_weights_kernel<<<B, 128, 0, as_stream(stream)>>>(y_id, row_of, T, g_row);
  HTCN_LAUNCH_CHECK("loss_row_weights_kernel");
  return HTCN_OK;
}

extern "C" int32_t htcn_adam_step(float* param, float* grad, float* m, float* v, int64_t n, float lr_t, float beta1,
                                  float beta2, float eps, const float* grad_div, int32_t zero_grad, void* stream) {
  HTCN_REQUIRE(param && grad && m && v && n > 0 && n % 4 == 0, "adam_step: bad args (n=%lld must be a multiple of 4)",
               (long long)n);
  adam_kernel<<<ceil_div(n / 4, 256), 256, 0, as_stream(stream)>>>(
      n / 4, reinterpret_cast<float4*>(param), reinterpret_cast<float4*>(grad), reinterpret_cast<float4*>(m),
      reinterpret_cast<float4*>(v), lr_t, beta1, beta2, eps, grad_div, zero_grad);
  HTCN_LAUNCH_CHECK("adam_kernel");
  return HTCN_OK;
}

extern "C" int32_t htcn_refresh_wout(const float* w_out_t_f32, const float* b_out, int32_t N, void* w_out_t,
                                     int32_t dtype, void* stream) {
  HTCN_REQUIRE(w_out_t_f32 && b_out && w_out_t && N > 0, "refresh_wout: bad args");
  HTCN_REQUIRE(dtype == HTCN_BF16, "refresh_wout: only the bf16 scoring layout is derived (the fp32 tier scores the master)");
  const long long n = (long long)N * (kWtPitchBf16 / 2);
  refresh_wout_bf16_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(w_out_t_f32, b_out, N,
                                                                           reinterpret_cast<__nv_bfloat16*>(w_out_t));
  HTCN_LAUNCH_CHECK("refresh_wout_bf16_kernel");
  return HTCN_OK;
}
