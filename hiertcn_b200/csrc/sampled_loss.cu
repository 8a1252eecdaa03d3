// Sampled ranking losses of reference loss.py:22-71 (nce / hinge_sigmoid / hinge_logsigmoid /
// hinge_linear / bpr).  pred [Q,128] is l2-normalised (tf.nn.l2_normalize: x * rsqrt(max(sum x^2, 1e-12)))
// and scored by inner product against the positive row and k negative rows of `table`.
// One warp per query row: 1 + k gathered 512 B rows (K1-style 128-bit loads), warp-shuffle dots.
#include "common.cuh"

namespace htcn {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float log_sigmoid(float x) {   // stable log(sigmoid(x))
  return fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// one lane's 4 channels of row `id` of the gather table: fp32 [N,128] rows (512 B), or -- kTabBf16 -- the bf16 scoring table
// [N,144] of htcn_prepare_wout (288 B rows, weights in columns 0..127): half the gathered bytes, the same values where the
// fp32 table is the widened bf16 one (the bf16 tier)
template <bool kTabBf16>
__device__ __forceinline__ float4 sl_load_row(const float4* __restrict__ table, long long id, int lane) {
  if (kTabBf16) {
    uint2 u;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
                 : "=r"(u.x), "=r"(u.y)
                 : "l"(reinterpret_cast<const uint8_t*>(table) + id * (HTCN_WT_PITCH_BF16 * 2) + lane * 8));
    return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
  }
  return ldg_nc_f4(table + id * 32 + lane);
}

template <bool kBf16, bool kTabBf16 = false>
__global__ void __launch_bounds__(256)
sampled_loss_kernel(const void* __restrict__ pred, int Q, const float4* __restrict__ table,
                    const int* __restrict__ pos_id, const int* __restrict__ neg_id, int k, int kind,
                    float delta, float nce_weight, int nce_div, float* __restrict__ loss_row) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  const int pos = pos_id[q];
  if (pos <= 0) {
    if (lane == 0) loss_row[q] = 0.f;
    return;
  }
  float4 p;
  if (kBf16) {
    const uint2 u = reinterpret_cast<const uint2*>(pred)[(long long)q * 32 + lane];
    p = make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
  } else {
    p = reinterpret_cast<const float4*>(pred)[(long long)q * 32 + lane];
  }
  const float ss = warp_sum(p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w);
  const float inv = rsqrtf(fmaxf(ss, 1e-12f));
  p.x *= inv; p.y *= inv; p.z *= inv; p.w *= inv;
  const float4 yp = sl_load_row<kTabBf16>(table, pos, lane);
  const float inner = warp_sum(p.x * yp.x + p.y * yp.y + p.z * yp.z + p.w * yp.w);
  float acc = 0.f;
  constexpr int kFly = 4;                                     // rows in flight per warp (8 with the bf16 table measured slower: 3.07 vs 2.58 ms)
  for (int j0 = 0; j0 < k; j0 += kFly) {
    float4 v[kFly];
    int id[kFly];
#pragma unroll
    for (int i = 0; i < kFly; ++i) {
      id[i] = (j0 + i < k) ? __ldg(neg_id + (long long)q * k + j0 + i) : -1;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (id[i] > 0) v[i] = sl_load_row<kTabBf16>(table, id[i], lane);   // id 0 -> zero row
    }
#pragma unroll
    for (int i = 0; i < kFly; ++i) {
      if (id[i] < 0) continue;
      const float s = warp_sum(p.x * v[i].x + p.y * v[i].y + p.z * v[i].z + p.w * v[i].w);
      switch (kind) {
        case HTCN_LOSS_NCE: acc += log_sigmoid(-s); break;                                           // :29
        case HTCN_LOSS_HINGE_SIGMOID: acc += fmaxf(sigmoidf_(s) - sigmoidf_(inner) + delta, 0.f); break;     // :37-40
        case HTCN_LOSS_HINGE_LOGSIGMOID: acc += fmaxf(log_sigmoid(s) - log_sigmoid(inner) + delta, 0.f); break;  // :47-50
        case HTCN_LOSS_HINGE_LINEAR: acc += fmaxf(s - inner + delta, 0.f); break;                    // :57-60
        case HTCN_LOSS_BPR: acc += log_sigmoid(sigmoidf_(inner) - sigmoidf_(s)); break;              // :65-70
      }
    }
  }
  if (lane == 0) {
    float out;
    if (kind == HTCN_LOSS_NCE) out = -log_sigmoid(inner) - acc / (float)nce_div * nce_weight;       // :31 (args.num_neg_sample)
    else if (kind == HTCN_LOSS_BPR) out = -acc / (float)k;
    else out = acc / (float)k;
    loss_row[q] = out;
  }
}

// Backward of the sampled ranking losses: given g[q] = dL/dloss_row[q], recomputes the scores of the row (one warp per
// query, the 1 + k rows are L2-hot from the forward) and forms
//     dL/ds+ and dL/ds-_j per loss kind (relu gradient where the hinge is strictly positive, like TF's ReluGrad),
//     d_phat = sum ds * y,   d_table[id] += ds * phat (float4 atomics; id 0 is the null item and owns no row),
//     d_pred = (d_phat - phat <phat, d_phat>) / ||p||      (tf.nn.l2_normalize; plain scale where ||p||^2 < 1e-12).
// k <= 32 (lane j keeps s-_j).
template <bool kBf16>
__global__ void __launch_bounds__(256)
sampled_loss_bwd_kernel(const void* __restrict__ pred, int Q, const float4* __restrict__ table,
                        const int* __restrict__ pos_id, const int* __restrict__ neg_id, int k, int kind, float delta,
                        float nce_weight, int nce_div, const float* __restrict__ g_row, float4* __restrict__ d_pred,
                        float* __restrict__ d_table) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  const int pos = pos_id[q];
  if (pos <= 0) {
    d_pred[(long long)q * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float4 p;
  if (kBf16) {
    const uint2 u = reinterpret_cast<const uint2*>(pred)[(long long)q * 32 + lane];
    p = make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
  } else {
    p = reinterpret_cast<const float4*>(pred)[(long long)q * 32 + lane];
  }
  const float ss = warp_sum(p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w);
  const float inv = rsqrtf(fmaxf(ss, 1e-12f));
  p.x *= inv; p.y *= inv; p.z *= inv; p.w *= inv;                     // phat
  const float4 yp = ldg_nc_f4(table + (long long)pos * 32 + lane);
  const float inner = warp_sum(p.x * yp.x + p.y * yp.y + p.z * yp.z + p.w * yp.w);
  // pass 1: s-_j -> lane j
  const int my_id = lane < k ? __ldg(neg_id + (long long)q * k + lane) : -1;
  float my_s = 0.f;
  for (int j = 0; j < k; ++j) {
    const int id = __shfl_sync(0xffffffffu, my_id, j);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (id > 0) v = ldg_nc_f4(table + (long long)id * 32 + lane);
    const float s = warp_sum(p.x * v.x + p.y * v.y + p.z * v.z + p.w * v.w);
    if (lane == j) my_s = s;
  }
  // dL/ds per lane (negatives) and for the positive
  const float g = g_row[q], fk = (float)k;
  float cj = 0.f, pos_part = 0.f;                                      // pos_part: this negative's share of dL/ds+
  if (lane < k) {
    const float s = my_s, sg = sigmoidf_(s), sgi = sigmoidf_(inner);
    switch (kind) {
      case HTCN_LOSS_NCE: cj = nce_weight / (float)nce_div * sg; break;
      case HTCN_LOSS_HINGE_SIGMOID: {
        const float a = (sg - sgi + delta > 0.f) ? 1.f : 0.f;
        cj = a * sg * (1.f - sg) / fk; pos_part = -a * sgi * (1.f - sgi) / fk; break;
      }
      case HTCN_LOSS_HINGE_LOGSIGMOID: {
        const float a = (log_sigmoid(s) - log_sigmoid(inner) + delta > 0.f) ? 1.f : 0.f;
        cj = a * sigmoidf_(-s) / fk; pos_part = -a * sigmoidf_(-inner) / fk; break;
      }
      case HTCN_LOSS_HINGE_LINEAR: {
        const float a = (s - inner + delta > 0.f) ? 1.f : 0.f;
        cj = a / fk; pos_part = -a / fk; break;
      }
      case HTCN_LOSS_BPR: {
        const float w = sigmoidf_(-(sgi - sg));                        // -d log(sigmoid(u)) / du, u = sigmoid(s+) - sigmoid(s-)
        cj = w * sg * (1.f - sg) / fk; pos_part = -w * sgi * (1.f - sgi) / fk; break;
      }
    }
  }
  float cpos = warp_sum(pos_part);
  if (kind == HTCN_LOSS_NCE) cpos = sigmoidf_(inner) - 1.f;
  cj *= g;
  cpos *= g;
  // pass 2: d_phat and the table rows
  float4 dph = make_float4(cpos * yp.x, cpos * yp.y, cpos * yp.z, cpos * yp.w);
  atomicAdd(reinterpret_cast<float4*>(d_table + (long long)pos * kDim + lane * 4),
            make_float4(cpos * p.x, cpos * p.y, cpos * p.z, cpos * p.w));
  for (int j = 0; j < k; ++j) {
    const int id = __shfl_sync(0xffffffffu, my_id, j);
    const float c = __shfl_sync(0xffffffffu, cj, j);
    if (id <= 0 || c == 0.f) continue;                                  // warp-uniform
    const float4 v = ldg_nc_f4(table + (long long)id * 32 + lane);
    dph.x = fmaf(c, v.x, dph.x); dph.y = fmaf(c, v.y, dph.y); dph.z = fmaf(c, v.z, dph.z); dph.w = fmaf(c, v.w, dph.w);
    atomicAdd(reinterpret_cast<float4*>(d_table + (long long)id * kDim + lane * 4), make_float4(c * p.x, c * p.y, c * p.z, c * p.w));
  }
  float4 out;
  if (ss >= 1e-12f) {
    const float dot = warp_sum(p.x * dph.x + p.y * dph.y + p.z * dph.z + p.w * dph.w);
    out = make_float4(inv * (dph.x - p.x * dot), inv * (dph.y - p.y * dot), inv * (dph.z - p.z * dot), inv * (dph.w - p.w * dot));
  } else {
    out = make_float4(inv * dph.x, inv * dph.y, inv * dph.z, inv * dph.w);
  }
  d_pred[(long long)q * 32 + lane] = out;
}

// calc_score (reference loss.py:76-105): score of k candidate rows per query, 'l2' = -||p - y'||^2 or 'inner_prod' = <p, y'>
// (pred is NOT normalised here, exactly like the reference).  One warp per query.
template <bool kBf16>
__global__ void __launch_bounds__(256)
calc_score_kernel(const void* __restrict__ pred, int Q, const float4* __restrict__ table, const int* __restrict__ cand_id,
                  int k, int mode, float* __restrict__ score) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  float4 p;
  if (kBf16) {
    const uint2 u = reinterpret_cast<const uint2*>(pred)[(long long)q * 32 + lane];
    p = make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
  } else {
    p = reinterpret_cast<const float4*>(pred)[(long long)q * 32 + lane];
  }
  for (int j0 = 0; j0 < k; j0 += 4) {
    float4 v[4];
    int id[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      id[i] = (j0 + i < k) ? __ldg(cand_id + (long long)q * k + j0 + i) : -1;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (id[i] > 0) v[i] = ldg_nc_f4(table + (long long)id[i] * 32 + lane);   // id 0 -> zero row
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (id[i] < 0) continue;
      float t;
      if (mode == 0) {
        const float dx = p.x - v[i].x, dy = p.y - v[i].y, dz = p.z - v[i].z, dw = p.w - v[i].w;
        t = -warp_sum(dx * dx + dy * dy + dz * dz + dw * dw);                  // loss.py:94
      } else {
        t = warp_sum(p.x * v[i].x + p.y * v[i].y + p.z * v[i].z + p.w * v[i].w);   // loss.py:96-97
      }
      if (lane == 0) score[(long long)q * k + j0 + i] = t;
    }
  }
}

}  // namespace htcn

extern "C" int32_t htcn_calc_score(const void* pred, int32_t precision, int32_t Q, const float* table,
                                   const int32_t* cand_id, int32_t k, int32_t rank_metric, float* score, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(pred && table && cand_id && score && Q > 0 && k > 0, "calc_score: bad args");
  HTCN_REQUIRE(rank_metric == 0 || rank_metric == 1, "calc_score: rank_metric %d (0 = l2, 1 = inner_prod)", rank_metric);
  const int grid = ceil_div(Q, 8);
  if (precision == HTCN_BF16)
    calc_score_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(pred, Q, (const float4*)table, cand_id, k, rank_metric, score);
  else if (precision == HTCN_F32)
    calc_score_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(pred, Q, (const float4*)table, cand_id, k, rank_metric, score);
  else
    HTCN_REQUIRE(false, "calc_score: precision %d", precision);
  HTCN_LAUNCH_CHECK("calc_score_kernel");
  return HTCN_OK;
}

// The same against the bf16 scoring table [N, HTCN_WT_PITCH_BF16] (htcn_prepare_wout, bf16 tier): pred must be bf16
extern "C" int32_t htcn_sampled_rank_loss_wt(const void* pred, int32_t Q, const void* wt_bf16, const int32_t* pos_id,
                                             const int32_t* neg_id, int32_t k, int32_t loss_kind, float hinge_delta,
                                             float nce_weight, int32_t num_neg_sample, float* loss_row, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(pred && wt_bf16 && pos_id && neg_id && loss_row && Q > 0 && k > 0, "sampled_rank_loss_wt: bad args");
  HTCN_REQUIRE(loss_kind >= HTCN_LOSS_NCE && loss_kind <= HTCN_LOSS_BPR, "sampled_rank_loss_wt: kind %d", loss_kind);
  const int nce_div = num_neg_sample > 0 ? num_neg_sample : k;
  sampled_loss_kernel<true, true><<<ceil_div(Q, 8), 256, 0, as_stream(stream)>>>(pred, Q, (const float4*)wt_bf16, pos_id, neg_id, k,
                                                                                loss_kind, hinge_delta, nce_weight, nce_div, loss_row);
  HTCN_LAUNCH_CHECK("sampled_loss_kernel");
  return HTCN_OK;
}

extern "C" int32_t htcn_sampled_rank_loss(const void* pred, int32_t precision, int32_t Q, const float* table,
                                          const int32_t* pos_id, const int32_t* neg_id, int32_t k,
                                          int32_t loss_kind, float hinge_delta, float nce_weight,
                                          int32_t num_neg_sample, float* loss_row, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(pred && table && pos_id && neg_id && loss_row && Q > 0 && k > 0, "sampled_rank_loss: bad args");
  HTCN_REQUIRE(loss_kind >= HTCN_LOSS_NCE && loss_kind <= HTCN_LOSS_BPR, "sampled_rank_loss: kind %d", loss_kind);
  const int nce_div = num_neg_sample > 0 ? num_neg_sample : k;
  const int grid = ceil_div(Q, 8);
  if (precision == HTCN_BF16)
    sampled_loss_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(pred, Q, (const float4*)table, pos_id, neg_id, k,
                                                                  loss_kind, hinge_delta, nce_weight, nce_div, loss_row);
  else if (precision == HTCN_F32)
    sampled_loss_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(pred, Q, (const float4*)table, pos_id, neg_id, k,
                                                                   loss_kind, hinge_delta, nce_weight, nce_div, loss_row);
  else
    HTCN_REQUIRE(false, "sampled_rank_loss: precision %d", precision);
  HTCN_LAUNCH_CHECK("sampled_loss_kernel");
  return HTCN_OK;
}


extern "C" int32_t htcn_sampled_rank_loss_backward(const void* pred, int32_t precision, int32_t Q, const float* table,
                                                   const int32_t* pos_id, const int32_t* neg_id, int32_t k,
                                                   int32_t loss_kind, float hinge_delta, float nce_weight,
                                                   int32_t num_neg_sample, const float* g_row, float* d_pred,
                                                   float* d_table, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(pred && table && pos_id && neg_id && g_row && d_pred && d_table && Q > 0, "sampled_rank_loss_backward: bad args");
  HTCN_REQUIRE(k > 0 && k <= 32, "sampled_rank_loss_backward: k=%d out of [1,32]", k);
  HTCN_REQUIRE(loss_kind >= HTCN_LOSS_NCE && loss_kind <= HTCN_LOSS_BPR, "sampled_rank_loss_backward: kind %d", loss_kind);
  const int nce_div = num_neg_sample > 0 ? num_neg_sample : k;
  const int grid = ceil_div(Q, 8);
  if (precision == HTCN_BF16)
    sampled_loss_bwd_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(pred, Q, (const float4*)table, pos_id, neg_id, k, loss_kind,
                                                                      hinge_delta, nce_weight, nce_div, g_row, (float4*)d_pred, d_table);
  else if (precision == HTCN_F32)
    sampled_loss_bwd_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(pred, Q, (const float4*)table, pos_id, neg_id, k, loss_kind,
                                                                       hinge_delta, nce_weight, nce_div, g_row, (float4*)d_pred, d_table);
  else
    HTCN_REQUIRE(false, "sampled_rank_loss_backward: precision %d", precision);
  HTCN_LAUNCH_CHECK("sampled_loss_bwd_kernel");
  return HTCN_OK;
}
