// K2 (fp32 tier): in-projection + dilated causal conv residual stack as FFMA tile GEMMs.
// One generic kernel evaluates one "level" for 128 consecutive flat rows x 128 output channels:
//     out[r,:] = epi( sum_{tap<K} in[r - (K-1-tap)*d, :] @ W[tap]  + bias [+ sbias[slot(r), user(r), :]] )
// with the shifted row taken as zero when it falls before the start of r's sequence -- the causal
// left pad of customized_tcn_cell.py:46-49 -- and epi = relu(relu(.) + res[r,:]) for conv levels
// (customized_tcn_cell.py:109-127: relu inside the conv, residual add, relu; res = in, or the 1x1 down-sample
// Dense(in) of :102-106,123-124 computed by a K = 1 launch of this kernel when the level changes the width) or
// identity for the in-projection (model_tcn.py:35, K = 1, no bias).  Levels are chained through an fp32 scratch in HBM
// (L2-resident at these sizes); the bf16 tier (k2_tcn_bf16.cu) keeps them on chip instead.
#include "train.cuh"

namespace htcn {

constexpr int kTM = 128;      // rows per CTA
constexpr int kTK = 16;       // contraction chunk
constexpr int kK2Threads = 256;


__device__ __forceinline__ float load_act(const void* base, bool bf16, long long idx) {
  if (bf16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
  return reinterpret_cast<const float*>(base)[idx];
}

__global__ void __launch_bounds__(kK2Threads)
k2_level_f32(LevelArgs a, SlotTable slots) {
  __shared__ __align__(16) float As[kTK][kTM];     // As[kk][row]
  __shared__ __align__(16) float Bs[kTK][kDim];    // Bs[kk][col]
  __shared__ int t_in_seq[kTM];                    // position of each tile row inside its sequence
  __shared__ int t_left[kTM];                      // rows remaining after it in the sequence (anti-causal taps)
  const int tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * kTM;
  // wide levels (129..256 channels = two 128-wide planes, fp32 tier only): this CTA computes output plane po
  const int po = blockIdx.y;
  const int n_in = a.in_planes > 1 ? a.in_planes : 1;
  if (po > 0 || n_in > 1) {
    a.w += (long long)po * n_in * a.K * kDim * kDim;
    if (a.bias) a.bias += po * kDim;
    a.out = reinterpret_cast<float*>(a.out) + po * a.out_plane_stride;          // wide paths are fp32 in / out
    if (a.resid) a.resid += po * a.resid_plane_stride;
  }

  if (tid < kTM) {
    const long long r = r0 + tid;
    int t = 0, left = 0;
    if (r < a.R) {
      const int p = (int)(r % a.T);
      int s = 0;
      while (s + 1 < slots.n && slots.off[s + 1] <= p) ++s;
      t = p - slots.off[s];
      left = slots.off[s + 1] - 1 - p;
    }
    t_in_seq[tid] = t;
    t_left[tid] = left;
  }
  __syncthreads();

  // thread tile: rows {ty*4..+3, 64+ty*4..+3} x cols {tx*4..+3, 64+tx*4..+3}
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int lrow = tid & 127, lhalf = tid >> 7;   // loader mapping: row, 8-wide k half
  for (int tap = 0; tap < n_in * a.K; ++tap) {
    const int shift = (a.K - 1 - tap % a.K) * a.dil;
    const long long src = (a.anti ? (r0 + lrow + shift) : (r0 + lrow - shift)) + (tap / a.K) * (a.in_plane_stride / kDim);
    const bool ok = (r0 + lrow < a.R) && (a.anti ? (shift <= t_left[lrow]) : (t_in_seq[lrow] - shift >= 0));
    for (int c0 = 0; c0 < kDim; c0 += kTK) {
      // A chunk: in[src][c0 + lhalf*8 .. +8] -> As[kk][lrow]
      float v[8];
      if (ok) {
        if (a.in_bf16) {
          const uint4 q = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.in) +
                                                          src * kDim + c0 + lhalf * 8);
          v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
          v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
        } else {
          const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.in) + src * kDim + c0 + lhalf * 8);
          const float4 q0 = p[0], q1 = p[1];
          v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      // W chunk: w[tap][c0 + kk][0..127] -> Bs[kk][col]; 2048 floats = 256 threads x 2 float4
      // (transposed use: Bs[kk][col] = w[tap][col][c0 + kk], each thread 8 consecutive kk of one col)
      const float4* wp0 = a.w_nt
          ? reinterpret_cast<const float4*>(a.w + ((long long)tap * kDim + lrow) * kDim + c0 + lhalf * 8)
          : reinterpret_cast<const float4*>(a.w + ((long long)tap * kDim + c0) * kDim) + tid;
      const float4 w0 = __ldg(wp0), w1 = __ldg(wp0 + (a.w_nt ? 1 : 256));
      __syncthreads();           // previous chunk fully consumed
#pragma unroll
      for (int i = 0; i < 8; ++i) As[lhalf * 8 + i][lrow] = v[i];
      if (a.w_nt) {
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) Bs[lhalf * 8 + i][lrow] = wv[i];
      } else {
        reinterpret_cast<float4*>(&Bs[0][0])[tid] = w0;
        reinterpret_cast<float4*>(&Bs[0][0])[256 + tid] = w1;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kTK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
  }

  // ---- epilogue ------------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int lr = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4);
    const long long r = r0 + lr;
    if (r >= a.R) continue;
    long long dst = r;
    if (a.out_row) {
      const int d = a.out_row[r];
      if (d < 0) continue;
      dst = d;
    }
    const float* sb = nullptr;
    const float* dr = nullptr;
    if (a.sbias || a.drop) {
      const int b = (int)(r / a.T), p = (int)(r % a.T);
      int s = 0;
      while (s + 1 < slots.n && slots.off[s + 1] <= p) ++s;
      if (a.sbias) sb = a.sbias + ((long long)s * a.B + b) * kDim;
      if (a.drop) dr = a.drop + (long long)s * a.drop_stride;
    }
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int c = jh * 64 + tx * 4;
      float o[4], ax[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[i][jh * 4 + j];
        if (a.bias) v += __ldg(a.bias + c + j);
        if (sb) v += __ldg(sb + c + j);
        if (a.conv_epilogue == 1 || a.conv_epilogue == 3) {
          v = fmaxf(v, 0.f);
          ax[j] = v;                                   // saved BEFORE the dropout scale (the backward re-applies it)
          if (dr) v *= __ldg(dr + c + j);
          const float res = a.conv_epilogue == 1 ? load_act(a.in, a.in_bf16, po * a.in_plane_stride + r * kDim + c + j)
                                                 : a.resid[r * kDim + c + j];
          v = fmaxf(v + res, 0.f);
        } else if (a.conv_epilogue == 2) {
          v += a.resid[r * kDim + c + j];
        }
        o[j] = v;
      }
      if (a.aux && (a.conv_epilogue == 1 || a.conv_epilogue == 3))
        *reinterpret_cast<float4*>(a.aux + r * kDim + c) = make_float4(ax[0], ax[1], ax[2], ax[3]);
      if (a.out_bf16) {
        uint2 q = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.out) + dst * kDim + c) = q;
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + dst * kDim + c) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

int32_t k2_level_launch(const LevelArgs& a, const SlotTable& slots, cudaStream_t st) {
  if (a.tc_ws && k2_level_tc_supported(a, slots)) return k2_level_tc_launch(a, slots, st);
  k2_level_f32<<<ceil_div(a.R, kTM), kK2Threads, 0, st>>>(a, slots);
  HTCN_LAUNCH_CHECK("k2_level_f32");
  return HTCN_OK;
}

// fp32 conv stack whose levels may be up to 256 channels wide (two 128-wide planes): activations are block-planar
// [P][R][128]; a level with P_in input planes runs as P_in*K taps, one CTA column per output plane (weights pre-arranged as
// [P_out, P_in*K, 128, 128] by hiertcn_b200.weights.to_device_layout).  hout: [P_last][n_out_rows][128] fp32.
int32_t tcn_forward_f32_wide(const void* xe, int xe_dtype, const float* w_in_x, const float* sbias,
                             const float* const* conv_w, const float* const* conv_b, const float* const* ds_w,
                             const float* const* ds_b, const int* planes, int n_levels, int K, const SlotTable& slots, int B,
                             int T, const int* out_row, float* hout, long long hout_plane_rows, float* scratch,
                             cudaStream_t st) {
  const long long R = (long long)B * T;
  const int grid = ceil_div(R, kTM);
  const long long plane = R * kDim;
  float* buf[2] = {scratch, scratch + 2 * plane};          // two planes each
  float* res_buf = scratch + 4 * plane;
  LevelArgs a{};
  a.R = R; a.T = T; a.B = B;
  a.in = xe; a.in_bf16 = (xe_dtype == HTCN_BF16);
  a.w = w_in_x; a.sbias = sbias; a.K = 1; a.dil = 1; a.conv_epilogue = 0;
  a.out = buf[0];
  k2_level_f32<<<grid, kK2Threads, 0, st>>>(a, slots);      // in-projection: always 128 wide
  HTCN_LAUNCH_CHECK("k2_level_f32(in-proj)");
  int p_in = 1;
  for (int l = 0; l < n_levels; ++l) {
    const bool last = (l == n_levels - 1);
    const int p_out = planes[l];
    const bool ds = ds_w && ds_w[l];
    if (!ds && p_in != p_out) {
      set_error("tcn_forward_wide: level %d changes the plane count without a down-sample kernel", l);
      return HTCN_ERR_INVALID;
    }
    a = LevelArgs{};
    a.R = R; a.T = T; a.B = B;
    a.in = buf[l & 1]; a.in_planes = p_in; a.in_plane_stride = plane;
    if (ds) {
      a.w = ds_w[l]; a.bias = ds_b ? ds_b[l] : nullptr; a.K = 1; a.dil = 1; a.conv_epilogue = 0;
      a.out = res_buf; a.out_plane_stride = plane;
      k2_level_f32<<<dim3(grid, p_out), kK2Threads, 0, st>>>(a, slots);
      HTCN_LAUNCH_CHECK("k2_level_f32(down-sample)");
    }
    a.w = conv_w[l]; a.bias = conv_b[l]; a.K = K; a.dil = 1 << l; a.conv_epilogue = ds ? 3 : 1;
    a.resid = ds ? res_buf : nullptr;
    a.resid_plane_stride = plane;
    a.out = last ? hout : buf[(l + 1) & 1];
    a.out_plane_stride = last ? hout_plane_rows * kDim : plane;
    a.out_row = last ? out_row : nullptr;
    k2_level_f32<<<dim3(grid, p_out), kK2Threads, 0, st>>>(a, slots);
    HTCN_LAUNCH_CHECK("k2_level_f32(conv)");
    p_in = p_out;
  }
  return HTCN_OK;
}

int32_t tcn_forward_f32(const void* xe, int xe_dtype, const float* w_in_x, const float* sbias,
                        const float* const* conv_w, const float* const* conv_b, const float* const* ds_w,
                        const float* const* ds_b, int n_levels, int K,
                        const SlotTable& slots, int B, int T, const int* out_row, void* hout, int hout_dtype,
                        float* scratch, cudaStream_t st) {
  const long long R = (long long)B * T;
  const int grid = ceil_div(R, kTM);
  float* buf[2] = {scratch, scratch + R * kDim};
  float* res_buf = scratch + 2 * R * kDim;        // down-sample residual of the current level (only with ds_w)
  LevelArgs a{};
  a.R = R; a.T = T; a.B = B;
  // in-projection: K = 1, no bias, per-sequence bias = state half of the concat (model_hier.py:54-55)
  a.in = xe; a.in_bf16 = (xe_dtype == HTCN_BF16);
  a.w = w_in_x; a.bias = nullptr; a.sbias = sbias; a.K = 1; a.dil = 1; a.conv_epilogue = 0;
  const bool last0 = (n_levels == 0);
  a.out = last0 ? hout : (void*)buf[0];
  a.out_row = last0 ? out_row : nullptr;
  a.out_bf16 = last0 ? (hout_dtype == HTCN_BF16) : 0;
  k2_level_f32<<<grid, kK2Threads, 0, st>>>(a, slots);
  HTCN_LAUNCH_CHECK("k2_level_f32(in-proj)");
  for (int l = 0; l < n_levels; ++l) {
    const bool last = (l == n_levels - 1);
    const bool ds = ds_w && ds_w[l];
    a.in = buf[l & 1]; a.in_bf16 = 0; a.sbias = nullptr;
    if (ds) {                                     // res = in @ W_ds + b_ds (customized_tcn_cell.py:102-106,123-124)
      a.w = ds_w[l]; a.bias = ds_b ? ds_b[l] : nullptr; a.K = 1; a.dil = 1; a.conv_epilogue = 0;
      a.out = res_buf; a.out_row = nullptr; a.out_bf16 = 0;
      k2_level_f32<<<grid, kK2Threads, 0, st>>>(a, slots);
      HTCN_LAUNCH_CHECK("k2_level_f32(down-sample)");
    }
    a.w = conv_w[l]; a.bias = conv_b[l]; a.K = K; a.dil = 1 << l; a.conv_epilogue = ds ? 3 : 1;
    a.resid = ds ? res_buf : nullptr;
    a.out = last ? hout : (void*)buf[(l + 1) & 1];
    a.out_row = last ? out_row : nullptr;
    a.out_bf16 = last ? (hout_dtype == HTCN_BF16) : 0;
    k2_level_f32<<<grid, kK2Threads, 0, st>>>(a, slots);
    HTCN_LAUNCH_CHECK("k2_level_f32(conv)");
  }
  return HTCN_OK;
}



}  // namespace htcn

extern "C" int32_t htcn_tcn_forward(const void* xe, int32_t xe_dtype, int32_t precision, const float* w_in_x,
                                    const float* sbias, const float* const* conv_w_host,
                                    const float* const* conv_b_host, const float* const* ds_w_host,
                                    const float* const* ds_b_host, int32_t n_levels, int32_t kernel_size,
                                    const int32_t* slot_off_host, int32_t B, int32_t T, int32_t S,
                                    const int32_t* out_row, void* hout, int32_t hout_dtype, float* scratch,
                                    void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(xe && w_in_x && hout && slot_off_host, "tcn_forward: NULL pointer");
  HTCN_REQUIRE(B > 0 && T > 0 && S > 0 && S <= HTCN_MAX_SLOTS, "tcn_forward: B=%d T=%d S=%d", B, T, S);
  HTCN_REQUIRE(n_levels >= 0 && n_levels <= HTCN_MAX_LEVELS, "tcn_forward: n_levels=%d", n_levels);
  HTCN_REQUIRE(kernel_size >= 1 && kernel_size <= 8, "tcn_forward: kernel_size=%d", kernel_size);
  HTCN_REQUIRE((xe_dtype == HTCN_F32 || xe_dtype == HTCN_BF16) && (hout_dtype == HTCN_F32 || hout_dtype == HTCN_BF16),
               "tcn_forward: bad dtype");
  HTCN_REQUIRE(n_levels == 0 || (conv_w_host && conv_b_host), "tcn_forward: conv weights NULL");
  SlotTable slots;
  slots.n = S;
  for (int i = 0; i <= S; ++i) slots.off[i] = slot_off_host[i];
  HTCN_REQUIRE(slots.off[0] == 0 && slots.off[S] == T, "tcn_forward: slot_off does not span T");
  if (precision == HTCN_F32) {
    HTCN_REQUIRE(scratch || n_levels == 0, "tcn_forward: f32 tier needs scratch");
    return tcn_forward_f32(xe, xe_dtype, w_in_x, sbias, conv_w_host, conv_b_host, ds_w_host, ds_b_host, n_levels, kernel_size, slots,
                           B, T, out_row, hout, hout_dtype, scratch, as_stream(stream));
  }
  if (precision == HTCN_BF16)
    return tcn_forward_bf16(xe, xe_dtype, w_in_x, sbias, conv_w_host, conv_b_host, ds_w_host, ds_b_host, n_levels, kernel_size, slots, B,
                            T, out_row, hout, hout_dtype, scratch, as_stream(stream));
  set_error("tcn_forward: precision %d", precision);
  return HTCN_ERR_INVALID;
}


extern "C" int32_t htcn_tcn_forward_wide(const void* xe, int32_t xe_dtype, const float* w_in_x, const float* sbias,
                                         const float* const* conv_w_host, const float* const* conv_b_host,
                                         const float* const* ds_w_host, const float* const* ds_b_host,
                                         const int32_t* level_planes_host, int32_t n_levels, int32_t kernel_size,
                                         const int32_t* slot_off_host, int32_t B, int32_t T, int32_t S,
                                         const int32_t* out_row, float* hout, int64_t hout_plane_rows, float* scratch,
                                         void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(xe && w_in_x && hout && slot_off_host && scratch && level_planes_host, "tcn_forward_wide: NULL pointer");
  HTCN_REQUIRE(B > 0 && T > 0 && S > 0 && S <= HTCN_MAX_SLOTS, "tcn_forward_wide: B=%d T=%d S=%d", B, T, S);
  HTCN_REQUIRE(n_levels >= 1 && n_levels <= HTCN_MAX_LEVELS && conv_w_host && conv_b_host, "tcn_forward_wide: n_levels=%d", n_levels);
  HTCN_REQUIRE(kernel_size >= 1 && kernel_size <= 8, "tcn_forward_wide: kernel_size=%d", kernel_size);
  HTCN_REQUIRE(xe_dtype == HTCN_F32 || xe_dtype == HTCN_BF16, "tcn_forward_wide: xe dtype");
  for (int l = 0; l < n_levels; ++l)
    HTCN_REQUIRE(level_planes_host[l] == 1 || level_planes_host[l] == 2, "tcn_forward_wide: level %d has %d planes", l, level_planes_host[l]);
  HTCN_REQUIRE(hout_plane_rows > 0, "tcn_forward_wide: hout_plane_rows");
  SlotTable slots;
  slots.n = S;
  for (int i = 0; i <= S; ++i) slots.off[i] = slot_off_host[i];
  HTCN_REQUIRE(slots.off[0] == 0 && slots.off[S] == T, "tcn_forward_wide: slot_off does not span T");
  return tcn_forward_f32_wide(xe, xe_dtype, w_in_x, sbias, conv_w_host, conv_b_host, ds_w_host, ds_b_host, level_planes_host,
                              n_levels, kernel_size, slots, B, T, out_row, hout, hout_plane_rows, scratch, as_stream(stream));
}
