// K2 training path (fp32): the conv stack forward with every level's activations saved, and its backward.
//
// forward (customized_tcn_cell.py:109-127, one conv per block):   p = conv_l(h_l) + b_l,  a_l = relu(p),
//                                                                 h_{l+1} = relu(a_l + h_l)
// backward, given dL/dh_{l+1}:   ds = dL/dh_{l+1} * [h_{l+1} > 0]        (TF ReluGrad: gradient where output > 0)
//                                dp = ds * [a_l > 0]
//                                dW_l[tap] += h_l[shifted by the tap]^T dp,   db_l += colsum(dp)
//                                dL/dh_l = ds + sum_tap dp[shifted the other way] W_l[tap]^T   (transposed convolution)
// a level with the 1x1 down-sample residual (h_{l+1} = relu(a_l + h_l Wds + bds), customized_tcn_cell.py:102-106,123-124):
//                                dWds += h_l^T ds,  dbds += colsum(ds),  dL/dh_l = ds Wds^T + (transposed convolution)
// and for the in-projection h_0 = Xe W_in[:128] + sbias[slot, user]  (model_hier.py:54-55, model_tcn.py:35):
//                                dW_in_x += Xe^T dh_0,  dsbias[s,b] = sum_{t in slot s} dh_0[b,t],  dXe = dh_0 W_in_x^T
#include "train.cuh"

namespace htcn {

namespace {

// drop (or NULL): [S][drop_stride] dropout scales of this level (forward: a_dropped = a * drop[slot, c]); T / slots locate
// the slot of a row
template <bool kBf16>
__global__ void relu_bwd_kernel(long long n4, float4* __restrict__ dcur, const void* __restrict__ h_next,
                                const void* __restrict__ a_l, float4* __restrict__ dp, const float* __restrict__ drop,
                                int drop_stride, int T, SlotTable slots) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 d = dcur[i];
  float4 h, a;
  if (kBf16) {
    const uint2 hq = reinterpret_cast<const uint2*>(h_next)[i], aq = reinterpret_cast<const uint2*>(a_l)[i];
    h = make_float4(bf16_lo(hq.x), bf16_hi(hq.x), bf16_lo(hq.y), bf16_hi(hq.y));
    a = make_float4(bf16_lo(aq.x), bf16_hi(aq.x), bf16_lo(aq.y), bf16_hi(aq.y));
  } else {
    h = reinterpret_cast<const float4*>(h_next)[i];
    a = reinterpret_cast<const float4*>(a_l)[i];
  }
  d.x = h.x > 0.f ? d.x : 0.f; d.y = h.y > 0.f ? d.y : 0.f; d.z = h.z > 0.f ? d.z : 0.f; d.w = h.w > 0.f ? d.w : 0.f;
  dcur[i] = d;
  float4 p;
  p.x = a.x > 0.f ? d.x : 0.f; p.y = a.y > 0.f ? d.y : 0.f; p.z = a.z > 0.f ? d.z : 0.f; p.w = a.w > 0.f ? d.w : 0.f;
  if (drop) {
    const int pos = (int)((i >> 5) % T);
    int s = 0;
    while (s + 1 < slots.n && slots.off[s + 1] <= pos) ++s;
    const float4 m = __ldg(reinterpret_cast<const float4*>(drop + (long long)s * drop_stride) + (i & 31));
    p.x *= m.x; p.y *= m.y; p.z *= m.z; p.w *= m.w;
  }
  dp[i] = p;
}

// dsbias[s, b, c] = sum_{t in slot s} dh0[b, t, c]
__global__ void slot_sum_kernel(const float* __restrict__ dh0, int B, int T, SlotTable slots, float* __restrict__ dsbias) {
  const int c = threadIdx.x;                   // 128 threads
  const int b = blockIdx.x, s = blockIdx.y;
  float acc = 0.f;
  for (int t = slots.off[s]; t < slots.off[s + 1]; ++t) acc += dh0[((long long)b * T + t) * kDim + c];
  dsbias[((long long)s * B + b) * kDim + c] = acc;
}

// dst[row_of[r], :] = src[r, :]  (gather = 1)   or   dst[r, :] = row_of[r] >= 0 ? src[row_of[r], :] : 0   (gather = 0)
__global__ void rows_compact_kernel(long long R, const int* __restrict__ row_of, const float4* __restrict__ src,
                                    float4* __restrict__ dst, int gather) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one float4
  if (i >= R * 32) return;
  const long long r = i >> 5;
  const int c = (int)(i & 31);
  const int q = row_of[r];
  if (gather) {
    if (q >= 0) dst[(long long)q * 32 + c] = src[i];
  } else {
    dst[i] = q >= 0 ? src[(long long)q * 32 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

}  // namespace
}  // namespace htcn

extern "C" int32_t htcn_tcn_forward_train(const float* xe, const float* w_in_x, const float* sbias,
                                          const float* const* conv_w_host, const float* const* conv_b_host,
                                          const float* const* ds_w_host, const float* const* ds_b_host,
                                          int32_t n_levels, int32_t kernel_size, const int32_t* slot_off_host, int32_t B,
                                          int32_t T, int32_t S, const int32_t* out_row, const float* dropout_scale,
                                          float* h_save, float* a_save, float* hout, void* tc_scratch, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(xe && w_in_x && h_save && hout && slot_off_host && out_row, "tcn_forward_train: NULL pointer");
  HTCN_REQUIRE(B > 0 && T > 0 && S > 0 && S <= HTCN_MAX_SLOTS, "tcn_forward_train: B=%d T=%d S=%d", B, T, S);
  HTCN_REQUIRE(n_levels >= 0 && n_levels <= HTCN_MAX_LEVELS && kernel_size >= 1 && kernel_size <= 8,
               "tcn_forward_train: n_levels=%d kernel_size=%d", n_levels, kernel_size);
  HTCN_REQUIRE(n_levels == 0 || (conv_w_host && conv_b_host && a_save), "tcn_forward_train: conv weights / a_save NULL");
  SlotTable slots;
  slots.n = S;
  for (int i = 0; i <= S; ++i) slots.off[i] = slot_off_host[i];
  HTCN_REQUIRE(slots.off[0] == 0 && slots.off[S] == T, "tcn_forward_train: slot_off does not span T");
  cudaStream_t st = as_stream(stream);
  const long long R = (long long)B * T;
  LevelArgs a{};
  a.R = R; a.T = T; a.B = B;
  // tc_scratch: every level on the tensor cores with fp32-grade split products (k2_level_tc.cu); NULL: the FFMA kernel
  a.tc_ws = tc_scratch; a.tc_split = 1;
  a.in = xe; a.w = w_in_x; a.sbias = sbias; a.K = 1; a.dil = 1; a.conv_epilogue = 0; a.out = h_save;
  int32_t rc = k2_level_launch(a, slots, st);
  if (rc) return rc;
  for (int l = 0; l < n_levels; ++l) {
    a.in = h_save + (long long)l * R * kDim;
    a.sbias = nullptr;
    const bool ds = ds_w_host && ds_w_host[l];
    if (ds) {
      // res = h_l Wds + bds, parked in a_save[l]: the conv launch below reads resid[r,c] and then writes aux[r,c] from the
      // same thread, so the two may share the buffer
      a.w = ds_w_host[l]; a.bias = ds_b_host ? ds_b_host[l] : nullptr; a.K = 1; a.dil = 1; a.conv_epilogue = 0;
      a.out = a_save + (long long)l * R * kDim; a.aux = nullptr; a.drop = nullptr;
      rc = k2_level_launch(a, slots, st);
      if (rc) return rc;
    }
    a.w = conv_w_host[l]; a.bias = conv_b_host[l]; a.K = kernel_size; a.dil = 1 << l;
    a.drop = dropout_scale ? dropout_scale + l * kDim : nullptr; a.drop_stride = n_levels * kDim;
    a.conv_epilogue = ds ? 3 : 1;
    a.resid = ds ? a_save + (long long)l * R * kDim : nullptr;
    a.out = h_save + (long long)(l + 1) * R * kDim;
    a.aux = a_save + (long long)l * R * kDim;
    rc = k2_level_launch(a, slots, st);
    if (rc) return rc;
  }
  rows_compact_kernel<<<ceil_div(R * 32, 256), 256, 0, st>>>(
      R, out_row, reinterpret_cast<const float4*>(h_save + (long long)n_levels * R * kDim),
      reinterpret_cast<float4*>(hout), 1);
  HTCN_LAUNCH_CHECK("rows_compact_kernel");
  return HTCN_OK;
}

extern "C" int32_t htcn_tcn_forward_train_bf16(const void* xe, const float* w_in_x, const float* sbias,
                                               const float* const* conv_w_host, const float* const* conv_b_host,
                                               const float* const* ds_w_host, const float* const* ds_b_host,
                                               int32_t n_levels, int32_t kernel_size, const int32_t* slot_off_host,
                                               int32_t B, int32_t T, int32_t S, const int32_t* out_row,
                                               const float* dropout_scale, void* h_save,
                                               void* a_save, void* hout, float* scratch, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(xe && w_in_x && h_save && hout && slot_off_host && out_row && scratch, "tcn_forward_train_bf16: NULL pointer");
  HTCN_REQUIRE(B > 0 && T > 0 && S > 0 && S <= HTCN_MAX_SLOTS, "tcn_forward_train_bf16: B=%d T=%d S=%d", B, T, S);
  HTCN_REQUIRE(n_levels >= 0 && n_levels <= HTCN_MAX_LEVELS && kernel_size >= 1 && kernel_size <= 8,
               "tcn_forward_train_bf16: n_levels=%d kernel_size=%d", n_levels, kernel_size);
  HTCN_REQUIRE(n_levels == 0 || (conv_w_host && conv_b_host && a_save), "tcn_forward_train_bf16: conv weights / a_save NULL");
  SlotTable slots;
  slots.n = S;
  for (int i = 0; i <= S; ++i) slots.off[i] = slot_off_host[i];
  HTCN_REQUIRE(slots.off[0] == 0 && slots.off[S] == T, "tcn_forward_train_bf16: slot_off does not span T");
  return tcn_forward_bf16(xe, HTCN_BF16, w_in_x, sbias, conv_w_host, conv_b_host, ds_w_host, ds_b_host, n_levels, kernel_size, slots, B, T, out_row,
                          hout, HTCN_BF16, scratch, as_stream(stream), h_save, a_save, dropout_scale);
}

extern "C" int32_t htcn_tcn_backward(const float* d_hout, const int32_t* out_row, const void* xe, int32_t save_dtype,
                                     const float* w_in_x, const float* const* conv_w_host,
                                     const float* const* ds_w_host, int32_t n_levels,
                                     int32_t kernel_size, const int32_t* slot_off_host, int32_t B, int32_t T, int32_t S,
                                     const float* dropout_scale, const void* h_save, const void* a_save, float* scratch,
                                     void* tc_scratch, float* const* d_conv_w_host, float* const* d_conv_b_host,
                                     float* const* d_ds_w_host, float* const* d_ds_b_host, float* d_w_in_x,
                                     float* d_sbias, float* d_xe, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(d_hout && out_row && xe && w_in_x && h_save && scratch && d_w_in_x && d_sbias && d_xe && slot_off_host,
               "tcn_backward: NULL pointer");
  HTCN_REQUIRE(save_dtype == HTCN_F32 || save_dtype == HTCN_BF16, "tcn_backward: save_dtype %d", save_dtype);
  HTCN_REQUIRE(B > 0 && T > 0 && S > 0 && S <= HTCN_MAX_SLOTS, "tcn_backward: B=%d T=%d S=%d", B, T, S);
  HTCN_REQUIRE(n_levels >= 0 && n_levels <= HTCN_MAX_LEVELS && kernel_size >= 1 && kernel_size <= 8,
               "tcn_backward: n_levels=%d kernel_size=%d", n_levels, kernel_size);
  HTCN_REQUIRE(n_levels == 0 || (conv_w_host && a_save && d_conv_w_host && d_conv_b_host), "tcn_backward: conv pointers NULL");
  SlotTable slots;
  slots.n = S;
  for (int i = 0; i <= S; ++i) slots.off[i] = slot_off_host[i];
  HTCN_REQUIRE(slots.off[0] == 0 && slots.off[S] == T, "tcn_backward: slot_off does not span T");
  cudaStream_t st = as_stream(stream);
  const bool bf = save_dtype == HTCN_BF16;
  const size_t esz = bf ? 2 : 4;
  const long long R = (long long)B * T;
  float* dcur = scratch;                    // [R,128] gradient flowing down the stack
  float* dp = scratch + R * kDim;           // [R,128] gradient at the conv pre-activation
  // tc_scratch != NULL: the weight gradients run on the tensor cores from zero-padded bf16 copies of the activations
  // (bwd_wgrad_bf16.cu, MN-major operands) instead of the fp32 split-K products
  const int P = n_levels > 0 ? (kernel_size - 1) * (1 << (n_levels - 1)) : 0;
  const PadGeom pg = make_pad_geom(slots, B, T, P);
  // layout: [bf16 weight tiles of the level being run (k2_level_tc.cu)][dp, zero-padded rows][h_l, zero-padded rows]
  __nv_bfloat16* bT = tc_scratch ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(tc_scratch) + HTCN_K2TC_WS_BYTES) : nullptr;
  __nv_bfloat16* aT = bT ? bT + 128 * pg.Kp : nullptr;
  const int zero = 0;
  const int eb = ceil_div(R * 32, 256);
  rows_compact_kernel<<<eb, 256, 0, st>>>(R, out_row, reinterpret_cast<const float4*>(d_hout),
                                          reinterpret_cast<float4*>(dcur), 0);
  HTCN_LAUNCH_CHECK("rows_compact_kernel(scatter)");
  int32_t rc;
  for (int l = n_levels - 1; l >= 0; --l) {
    const uint8_t* h_l = reinterpret_cast<const uint8_t*>(h_save) + (size_t)l * R * kDim * esz;
    const uint8_t* h_n = reinterpret_cast<const uint8_t*>(h_save) + (size_t)(l + 1) * R * kDim * esz;
    const uint8_t* a_l = reinterpret_cast<const uint8_t*>(a_save) + (size_t)l * R * kDim * esz;
    if (bf)
      relu_bwd_kernel<true><<<eb, 256, 0, st>>>(R * 32, reinterpret_cast<float4*>(dcur), h_n, a_l, reinterpret_cast<float4*>(dp),
                                                dropout_scale ? dropout_scale + l * kDim : nullptr, n_levels * kDim, T, slots);
    else
      relu_bwd_kernel<false><<<eb, 256, 0, st>>>(R * 32, reinterpret_cast<float4*>(dcur), h_n, a_l, reinterpret_cast<float4*>(dp),
                                                 dropout_scale ? dropout_scale + l * kDim : nullptr, n_levels * kDim, T, slots);
    HTCN_LAUNCH_CHECK("relu_bwd_kernel");
    const int dil = 1 << l;
    if (aT) {
      int shifts[8];                            // MN-major operands: one padded copy each, the tap shift is a TMA row offset
      for (int tap = 0; tap < kernel_size; ++tap) shifts[tap] = (kernel_size - 1 - tap) * dil;
      rc = pad_rows_bf16(h_l, bf, pg, aT, st);
      if (rc) return rc;
      rc = pad_rows_bf16(dp, false, pg, bT, st);
      if (rc) return rc;
      rc = wgrad_mn_bf16(aT, bT, pg.Kp, shifts, kernel_size, d_conv_w_host[l], st);
      if (rc) return rc;
    } else {
      for (int tap = 0; tap < kernel_size; ++tap) {
        rc = sgemm_tn_atomic(R, h_l, kDim, dp, kDim, d_conv_w_host[l] + (long long)tap * kDim * kDim, kDim,
                             (kernel_size - 1 - tap) * dil, T, &slots, st, bf);
        if (rc) return rc;
      }
    }
    rc = colsum_atomic(R, dp, kDim, kDim, d_conv_b_host[l], st);
    if (rc) return rc;
    const float* resid = dcur;                  // identity residual: dL/dh_l = ds + ...
    if (ds_w_host && ds_w_host[l]) {            // down-sample residual: dWds += h_l^T ds, dbds += colsum(ds), ds Wds^T + ...
      HTCN_REQUIRE(d_ds_w_host && d_ds_w_host[l] && d_ds_b_host && d_ds_b_host[l], "tcn_backward: down-sample gradient pointers NULL");
      if (aT) {                                  // aT still holds the padded h_l
        rc = pad_rows_bf16(dcur, false, pg, bT, st);
        if (rc) return rc;
        rc = wgrad_mn_bf16(aT, bT, pg.Kp, &zero, 1, d_ds_w_host[l], st);
      } else {
        rc = sgemm_tn_atomic(R, h_l, kDim, dcur, kDim, d_ds_w_host[l], kDim, 0, T, nullptr, st, bf);
      }
      if (rc) return rc;
      rc = colsum_atomic(R, dcur, kDim, kDim, d_ds_b_host[l], st);
      if (rc) return rc;
      float* dres = scratch + 2 * R * kDim;     // third scratch plane
      LevelArgs d{};                            // dres = ds Wds^T
      d.R = R; d.T = T; d.B = B; d.tc_ws = tc_scratch;
      d.in = dcur; d.w = ds_w_host[l]; d.K = 1; d.dil = 1; d.conv_epilogue = 0; d.out = dres; d.anti = 1; d.w_nt = 1;
      if (tc_scratch && k2_level_tc_supported(d, slots)) rc = k2_level_tc_launch(d, slots, st);
      else rc = sgemm(true, R, kDim, kDim, dcur, kDim, ds_w_host[l], kDim, dres, kDim, false, st);
      if (rc) return rc;
      resid = dres;
    }
    LevelArgs a{};
    a.R = R; a.T = T; a.B = B; a.tc_ws = tc_scratch;      // plain bf16 products: the gates are the saved forward's
    a.in = dp; a.w = conv_w_host[l]; a.K = kernel_size; a.dil = dil; a.conv_epilogue = 2; a.resid = resid; a.out = dcur;
    a.anti = 1; a.w_nt = 1;
    rc = k2_level_launch(a, slots, st);
    if (rc) return rc;
  }
  if (aT) {
    rc = pad_rows_bf16(xe, bf, pg, aT, st);
    if (rc) return rc;
    rc = pad_rows_bf16(dcur, false, pg, bT, st);
    if (rc) return rc;
    rc = wgrad_mn_bf16(aT, bT, pg.Kp, &zero, 1, d_w_in_x, st);
  } else {
    rc = sgemm_tn_atomic(R, xe, kDim, dcur, kDim, d_w_in_x, kDim, 0, T, nullptr, st, bf);
  }
  if (rc) return rc;
  slot_sum_kernel<<<dim3(B, S), kDim, 0, st>>>(dcur, B, T, slots, d_sbias);
  HTCN_LAUNCH_CHECK("slot_sum_kernel");
  LevelArgs d{};                                // dXe = dh_0 W_in_x^T
  d.R = R; d.T = T; d.B = B; d.tc_ws = tc_scratch;
  d.in = dcur; d.w = w_in_x; d.K = 1; d.dil = 1; d.conv_epilogue = 0; d.out = d_xe; d.anti = 1; d.w_nt = 1;
  if (tc_scratch && k2_level_tc_supported(d, slots)) return k2_level_tc_launch(d, slots, st);
  return sgemm(true, R, kDim, kDim, dcur, kDim, w_in_x, kDim, d_xe, kDim, false, st);
}


extern "C" int64_t htcn_tcn_backward_tc_scratch_bytes(int32_t B, int32_t T, int32_t S, int32_t n_levels, int32_t kernel_size) {
  const long long P = n_levels > 0 ? (long long)(kernel_size - 1) * (1 << (n_levels - 1)) : 0;
  const long long Kp = ((long long)B * (T + (long long)S * P) + 63) / 64 * 64;
  (void)kernel_size;                          // level weight tiles + zero-padded bf16 copies of dp and of h_l
  return htcn::HTCN_K2TC_WS_BYTES + 2 * 128 * Kp * 2;
}
