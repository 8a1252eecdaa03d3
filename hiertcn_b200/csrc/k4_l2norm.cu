// l2-normalised scoring head (reference model_tcn.py:42-43: pred = tf.nn.l2_normalize(pred, dim=-1) on the [B,L,N] logits)
// without materialising the logits.  The squared norm of a logits row is a quadratic form of the user embedding:
//     sum_j (h . w_j + b_j)^2 = h^T G h + 2 h . c + s,     G = sum_j w_j w_j^T [128,128],  c = sum_j b_j w_j,  s = sum_j b_j^2
// G, c, s depend on the output table only (htcn_catalog_gram, once per weight update); the per-row scale
// 1 / sqrt(max(||z||^2, eps)) (htcn_logit_rownorm) then rides through the CE sweep as a per-row multiplier of the exponent
// (htcn_score_ce_rank_l2norm).  Strict-greater ranks and top-k ORDER are invariant under the positive scale; top-k VALUES are
// scaled afterwards (htcn_scale_rows).
#include "common.cuh"

namespace htcn {
int32_t score_f32(const ScoreArgs& a, cudaStream_t st);
int32_t score_bf16(const ScoreArgs& a, cudaStream_t st);

namespace {
constexpr int kGramBlocks = 148;
constexpr int kGramFloats = kDim * kDim + kDim + 4;       // G | c | s, pad

// per-block partial of G / c / s over a contiguous range of catalog rows; thread (ty, tx) owns the 8x8 tile
// G[ty*8.., tx*8..]; deterministic: fixed ranges, fixed order, partials summed by gram_reduce_kernel
template <bool kBf16>
__global__ void __launch_bounds__(256)
gram_partial_kernel(const void* __restrict__ wt, const float* __restrict__ b_out, int n_items, float* __restrict__ part) {
  __shared__ float row[8][kDim + 1];                      // 8 rows per step; [.. ][128] = the row's bias
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long per = (n_items + gridDim.x - 1) / gridDim.x;
  const long long j0 = (long long)blockIdx.x * per, j1 = min((long long)n_items, j0 + per);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float cacc = 0.f, sacc = 0.f;                           // threads 0..127: c[tid]; thread 0: s
  for (long long base = j0; base < j1; base += 8) {
    __syncthreads();
    for (int e = tid; e < 8 * (kDim + 1); e += 256) {
      const int r = e / (kDim + 1), k = e % (kDim + 1);
      const long long j = base + r;
      float v = 0.f;
      if (j < j1) {
        if (kBf16) {
          const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(wt) + j * kWtPitchBf16;
          v = k < kDim ? __bfloat162float(p[k]) : __bfloat162float(p[kDim]) + __bfloat162float(p[kDim + 1]);
        } else {
          v = k < kDim ? reinterpret_cast<const float*>(wt)[j * kDim + k] : b_out[j];
        }
      }
      row[r][k] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] = row[r][ty * 8 + i];
        b[i] = row[r][tx * 8 + i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      const float bj = row[r][kDim];
      if (tid < kDim) cacc = fmaf(bj, row[r][tid], cacc);
      if (tid == 0) sacc = fmaf(bj, bj, sacc);
    }
  }
  float* out = part + (long long)blockIdx.x * kGramFloats;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) out[(ty * 8 + i) * kDim + tx * 8 + j] = acc[i][j];
  if (tid < kDim) out[kDim * kDim + tid] = cacc;
  if (tid == 0) out[kDim * kDim + kDim] = sacc;
}

__global__ void gram_reduce_kernel(const float* __restrict__ part, int n_part, float* __restrict__ gram) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > kDim * kDim + kDim) return;
  double s = 0.0;
  for (int p = 0; p < n_part; ++p) s += (double)part[(long long)p * kGramFloats + i];
  gram[i] = (float)s;
}

// scale[q] = rsqrt(max(h^T G h + 2 h.c + s, eps)); one warp per row, G in shared memory (padded rows: conflict-free)
template <bool kBf16>
__global__ void __launch_bounds__(256)
rownorm_kernel(const void* __restrict__ hout, int Q, const float* __restrict__ gram, float eps, float* __restrict__ scale) {
  extern __shared__ float gs[];                           // [128][129] G, [128] c
  float* cs = gs + kDim * (kDim + 1);
  for (int e = threadIdx.x; e < kDim * kDim; e += 256) gs[(e / kDim) * (kDim + 1) + e % kDim] = gram[e];
  for (int e = threadIdx.x; e < kDim; e += 256) cs[e] = gram[kDim * kDim + e];
  __syncthreads();
  const float sb = gram[kDim * kDim + kDim];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long q = (long long)blockIdx.x * 8 + warp; q < Q; q += (long long)gridDim.x * 8) {
    float h[4];                                           // lane owns k = lane, lane+32, lane+64, lane+96
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long idx = q * kDim + lane + 32 * u;
      h[u] = kBf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(hout)[idx]) : reinterpret_cast<const float*>(hout)[idx];
    }
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) acc = fmaf(2.f * h[u], cs[lane + 32 * u], acc);
    for (int k = 0; k < kDim; ++k) {                      // (G h)_row for the lane's 4 rows needs every h_k: broadcast
      const float hk = __shfl_sync(0xffffffffu, h[k >> 5], k & 31);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = fmaf(h[u] * hk, gs[(lane + 32 * u) * (kDim + 1) + k], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) scale[q] = rsqrtf(fmaxf(acc + sb, eps));
  }
}

__global__ void scale_rows_kernel(float* __restrict__ x, const float* __restrict__ scale, long long R, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R * C) x[i] *= scale[i / C];
}

__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * b[i];
}
}  // namespace
}  // namespace htcn

extern "C" int64_t htcn_catalog_gram_scratch_floats(void) { return (int64_t)htcn::kGramBlocks * htcn::kGramFloats; }

extern "C" int32_t htcn_catalog_gram(const void* w_out_t, int32_t precision, const float* b_out, int32_t n_items,
                                     float* scratch, float* gram, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(w_out_t && scratch && gram && n_items > 0, "catalog_gram: bad args");
  HTCN_REQUIRE(precision == HTCN_BF16 || (precision == HTCN_F32 && b_out), "catalog_gram: precision %d / b_out", precision);
  cudaStream_t st = as_stream(stream);
  if (precision == HTCN_BF16) gram_partial_kernel<true><<<kGramBlocks, 256, 0, st>>>(w_out_t, b_out, n_items, scratch);
  else gram_partial_kernel<false><<<kGramBlocks, 256, 0, st>>>(w_out_t, b_out, n_items, scratch);
  HTCN_LAUNCH_CHECK("gram_partial_kernel");
  gram_reduce_kernel<<<ceil_div(kDim * kDim + kDim + 1, 256), 256, 0, st>>>(scratch, kGramBlocks, gram);
  HTCN_LAUNCH_CHECK("gram_reduce_kernel");
  return HTCN_OK;
}

extern "C" int32_t htcn_logit_rownorm(const void* hout, int32_t precision, int32_t Q, const float* gram, float eps,
                                      float* row_scale, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(hout && gram && row_scale && Q > 0, "logit_rownorm: bad args");
  HTCN_REQUIRE(precision == HTCN_F32 || precision == HTCN_BF16, "logit_rownorm: precision %d", precision);
  const size_t smem = sizeof(float) * (kDim * (kDim + 1) + kDim);
  const int grid = ceil_div(Q, 8) < 148 * 2 ? ceil_div(Q, 8) : 148 * 2;
  if (precision == HTCN_BF16) {
    HTCN_CUDA(cudaFuncSetAttribute(rownorm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rownorm_kernel<true><<<grid, 256, smem, as_stream(stream)>>>(hout, Q, gram, eps, row_scale);
  } else {
    HTCN_CUDA(cudaFuncSetAttribute(rownorm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rownorm_kernel<false><<<grid, 256, smem, as_stream(stream)>>>(hout, Q, gram, eps, row_scale);
  }
  HTCN_LAUNCH_CHECK("rownorm_kernel");
  return HTCN_OK;
}

extern "C" int32_t htcn_scale_rows(float* x, const float* row_scale, int64_t R, int32_t C, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(x && row_scale && R >= 0 && C > 0, "scale_rows: bad args");
  if (R == 0) return HTCN_OK;
  scale_rows_kernel<<<ceil_div(R * C, 256), 256, 0, as_stream(stream)>>>(x, row_scale, R, C);
  HTCN_LAUNCH_CHECK("scale_rows_kernel");
  return HTCN_OK;
}

extern "C" int32_t htcn_score_ce_rank_l2norm(const void* hout, int32_t precision, int32_t Q, const void* w_out_t,
                                             const float* b_out, int32_t n_items, int32_t n0, const int32_t* y_id,
                                             const float* target_logit, const float* row_scale, int32_t n_split,
                                             float* part_max, float* part_sum, int32_t* part_cnt, float* target_scaled,
                                             void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(hout && w_out_t && y_id && target_logit && row_scale && part_max && part_sum && part_cnt && target_scaled,
               "score_l2norm: NULL pointer");
  HTCN_REQUIRE(Q > 0 && n_items > 0 && n_split >= 1, "score_l2norm: Q=%d n_items=%d n_split=%d", Q, n_items, n_split);
  HTCN_REQUIRE(b_out || precision == HTCN_BF16, "score_l2norm: b_out is NULL");
  cudaStream_t st = as_stream(stream);
  mul_kernel<<<ceil_div(Q, 256), 256, 0, st>>>(target_logit, row_scale, target_scaled, Q);
  HTCN_LAUNCH_CHECK("mul_kernel");
  ScoreArgs a{hout, w_out_t, b_out, y_id, target_logit, part_max, part_sum, part_cnt, nullptr, nullptr,
              Q, n_items, n0, 0, n_split, HTCN_SCORE_CE | HTCN_SCORE_RANK};
  a.row_scale = row_scale;
  if (precision == HTCN_F32) return score_f32(a, st);
  if (precision == HTCN_BF16) return score_bf16(a, st);
  HTCN_REQUIRE(false, "score_l2norm: precision %d", precision);
}
