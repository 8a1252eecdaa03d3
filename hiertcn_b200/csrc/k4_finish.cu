// K4 entry points (dispatch to the fp32 FFMA tier or the bf16 tcgen05 tier) and the small
// finishing kernels: W_out layout preparation, partial merge (CE / rank), top-k merge, and the
// masked two-level means of model.py:111-117 / loss.py:190-219.
#include "common.cuh"

namespace htcn {

int32_t score_f32(const ScoreArgs& a, cudaStream_t st);
int32_t target_logit_f32(const float* hout, const float* wt, const float* b_out, const int* y_id, int Q,
                         int n_items, int n0, float* zy, cudaStream_t st, int planes = 1);
static inline bool is_f32(int precision) { return precision == HTCN_F32 || precision == HTCN_F32_W256; }
static inline int planes_of(int precision) { return precision == HTCN_F32_W256 ? 2 : 1; }
// bf16 tier (k4_score_bf16.cu)
int32_t score_bf16(const ScoreArgs& a, cudaStream_t st);
int32_t score_ce_rank_folded_bf16(const ScoreArgs& a, void* workspace, cudaStream_t st);
int32_t target_logit_bf16(const void* hout, const void* wt, const float* b_out, const int* y_id, int Q,
                          int n_items, int n0, float* zy, cudaStream_t st);
long long topk_workspace_bytes_bf16(int Q, int n_items, int k, int n_split);
int32_t score_topk_bf16(const void* hout, int Q, const void* wt, int n_items, int n0, int k, int n_split, void* workspace,
                        long long workspace_bytes, float* out_val, int* out_idx, int* overflow_rows, cudaStream_t st,
                        const ScoreArgs* ce = nullptr);

// the two-pass method needs >= k finite column groups of 256 items per row to produce a threshold
static bool topk_two_pass_ok(int precision, int n_items, int k) { return precision == HTCN_BF16 && n_items >= 1024 * k; }
static int topk_heap_splits(int precision, int n_items, int n_split) {
  const int tiles = precision == HTCN_BF16 ? (n_items + 127) / 128 : (n_items + 63) / 64;
  return n_split < 1 ? 1 : (n_split > tiles ? tiles : n_split);
}

// ---- W_out [128, N] f32 -> W_out^T [N, 128] f32 | [N, 144] bf16 augmented (32x32 smem transpose) ----
template <bool kBf16>
__global__ void prepare_wout_kernel(const float* __restrict__ w, const float* __restrict__ b, int N,
                                    void* __restrict__ out, int f32_pitch = kDim) {
  __shared__ float tile[32][33];
  const int pitch = kBf16 ? kWtPitchBf16 : f32_pitch;     // f32: 128, or 256 for HTCN_F32_W256 (grid.y = pitch / 32)
  const int n0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  if (kBf16 && blockIdx.y == kDim / 32) {         // bf16 only: the bias columns 128..143 of 32 items
    if (ty == 0 && n0 + tx < N) {
      __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(out) + (long long)(n0 + tx) * pitch + kDim;
      const float bv = b[n0 + tx];
      const __nv_bfloat16 hi = __float2bfloat16_rn(bv);
      const __nv_bfloat16 lo = __float2bfloat16_rn(bv - __bfloat162float(hi));
      // [b_hi, b_lo | b_hi, b_lo, 1, 1, 1 | 0 x 9]: the sweeps multiply the first pair by A's constant [1, 1]; the folded
      // CE + rank sweep (kFold, k4_score_bf16.cu) uses all seven: (c_hi + c_lo)(b_hi + b_lo) + t1 + t2 + t3
      row[0] = hi; row[1] = lo; row[2] = hi; row[3] = lo;
      row[4] = row[5] = row[6] = __float2bfloat16_rn(1.f);
      for (int i = 7; i < pitch - kDim; ++i) row[i] = __float2bfloat16_rn(0.f);
    }
    return;
  }
  const int c0 = blockIdx.y * 32;
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + tx;
    tile[i][tx] = (n < N) ? w[(long long)(c0 + i) * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + i;
    if (n < N) {
      const float v = tile[tx][i];
      if (kBf16) reinterpret_cast<__nv_bfloat16*>(out)[(long long)n * pitch + c0 + tx] = __float2bfloat16_rn(v);
      else reinterpret_cast<float*>(out)[(long long)n * pitch + c0 + tx] = v;
    }
  }
}

// ---- full logits for small catalogs (the tensor the reference materialises) ---------------------
__global__ void score_logits_kernel(const void* __restrict__ hout, int h_bf16, int Q, const void* __restrict__ wt,
                                    int w_bf16, const float* __restrict__ b_out, int n_items,
                                    float* __restrict__ logits, int planes) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Q * n_items) return;
  const int q = (int)(i / n_items), j = (int)(i % n_items);
  const __nv_bfloat16* wb = reinterpret_cast<const __nv_bfloat16*>(wt) + (long long)j * kWtPitchBf16;
  float acc = 0.f;
  for (int p = 0; p < planes; ++p)            // planes > 1: fp32 only, hout block-planar [P][Q][128], wt rows P*128 floats
    for (int k = 0; k < kDim; ++k) {
      const float a = h_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(hout)[(long long)q * kDim + k])
                             : reinterpret_cast<const float*>(hout)[((long long)p * Q + q) * kDim + k];
      const float b = w_bf16 ? __bfloat162float(wb[k]) : reinterpret_cast<const float*>(wt)[((long long)j * planes + p) * kDim + k];
      acc = fmaf(a, b, acc);
    }
  // bf16 tier: the bias is the (hi, lo) pair stored in the augmented columns, as the tensor-core sweep sees it
  const float bias = w_bf16 ? (__bfloat162float(wb[kDim]) + __bfloat162float(wb[kDim + 1])) : b_out[j];
  logits[i] = acc + bias;
}

// ---- merge CE / rank partials -----------------------------------------------------------------
__global__ void score_finish_kernel(const float* __restrict__ pm, const float* __restrict__ ps,
                                    const int* __restrict__ pc, int n_part, int Q, const int* __restrict__ y_id,
                                    const float* __restrict__ zy, float* __restrict__ loss_row,
                                    float* __restrict__ rank_row) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  if (loss_row) {
    float M = -INFINITY;
    for (int p = 0; p < n_part; ++p) M = fmaxf(M, pm[(long long)p * Q + q]);
    float s = 0.f;
    for (int p = 0; p < n_part; ++p) s += ps[(long long)p * Q + q] * expf(pm[(long long)p * Q + q] - M);
    const bool valid = !y_id || y_id[q] > 0;       // id 0: the label row is all zero -> loss 0
    loss_row[q] = valid ? (M + logf(s)) - zy[q] : 0.f;
  }
  if (rank_row) {
    int c = 0;
    for (int p = 0; p < n_part; ++p) c += pc[(long long)p * Q + q];
    rank_row[q] = (float)c;
  }
}

// ---- exact redo of CE rows whose tensor-core partial sum overflowed ---------------------------------------------------
// The bf16 sweep keeps sum_j 2^((z_j - z_y) log2e) with the TARGET logit as reference point (no running max): a row whose
// best logit exceeds its target logit by more than ~88 overflows fp32 and comes out of score_finish as +inf / NaN.  Such
// rows (softmax probability of the target < e^-88) are redone here with an online-max log-sum-exp over the whole
// catalog, one CTA per row, same bf16 operands; rows with a finite loss are skipped (one load each).
// Shard mode (loss_row == NULL; catalog-sharded scoring, hiertcn_b200/dist.py): a row is bad when one of THIS shard's
// n_split partial sums is non-finite; its exact (max, sum exp(z - max)) over the shard replaces the shard's partials
// (split 0 gets the pair, the other splits the neutral element (-inf, 0)), so that the cross-shard log-sum-exp merge of
// htcn_score_finish -- which accepts any reference point per part -- comes out finite.
__global__ void __launch_bounds__(256)
ce_repair_bf16_kernel(const __nv_bfloat16* __restrict__ hout, int Q, const __nv_bfloat16* __restrict__ wt, int n_items,
                      const float* __restrict__ zy, float* __restrict__ loss_row, float* __restrict__ part_max,
                      float* __restrict__ part_sum, int n_split, int* __restrict__ repaired) {
  __shared__ float h[kDim];
  __shared__ float red_m[8], red_s[8];
  __shared__ int bad_rows[256];
  __shared__ int n_bad;
  for (int base = blockIdx.x * 256; base < Q; base += gridDim.x * 256) {       // 256 rows checked per pass, one per thread
    const int qc = base + threadIdx.x;
    bool bad = false;
    if (qc < Q) {
      if (loss_row) bad = !isfinite(loss_row[qc]);
      else
        for (int p = 0; p < n_split; ++p) bad |= !isfinite(part_sum[(long long)p * Q + qc]);
    }
    if (!__syncthreads_or(bad)) continue;                         // the common case: nothing to redo in this chunk
    if (threadIdx.x == 0) n_bad = 0;
    __syncthreads();
    if (bad) bad_rows[atomicAdd(&n_bad, 1)] = qc;
    __syncthreads();
    const int nb = n_bad;
    for (int bi = 0; bi < nb; ++bi) {
    const int q = bad_rows[bi];
    __syncthreads();
    if (threadIdx.x < kDim) h[threadIdx.x] = __bfloat162float(hout[(long long)q * kDim + threadIdx.x]);
    __syncthreads();
    float m = -INFINITY, sum = 0.f;
    for (int j = threadIdx.x; j < n_items; j += blockDim.x) {
      const uint4* row = reinterpret_cast<const uint4*>(wt + (long long)j * kWtPitchBf16);
      float z = 0.f;
#pragma unroll 4
      for (int c = 0; c < kDim / 8; ++c) {
        const uint4 w = __ldg(row + c);
        const float* hh = h + c * 8;
        z = fmaf(hh[0], bf16_lo(w.x), z); z = fmaf(hh[1], bf16_hi(w.x), z);
        z = fmaf(hh[2], bf16_lo(w.y), z); z = fmaf(hh[3], bf16_hi(w.y), z);
        z = fmaf(hh[4], bf16_lo(w.z), z); z = fmaf(hh[5], bf16_hi(w.z), z);
        z = fmaf(hh[6], bf16_lo(w.w), z); z = fmaf(hh[7], bf16_hi(w.w), z);
      }
      const uint32_t bb = *reinterpret_cast<const uint32_t*>(wt + (long long)j * kWtPitchBf16 + kDim);   // b_hi | b_lo
      z += bf16_lo(bb) + bf16_hi(bb);
      const float mn = fmaxf(m, z);
      sum = sum * __expf(m - mn) + __expf(z - mn);
      m = mn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, sum, o);
      const float mn = fmaxf(m, m2);
      sum = (mn == -INFINITY) ? 0.f : sum * __expf(m - mn) + s2 * __expf(m2 - mn);
      m = mn;
    }
    if ((threadIdx.x & 31) == 0) { red_m[threadIdx.x >> 5] = m; red_s[threadIdx.x >> 5] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float M = -INFINITY;
      for (int w = 0; w < 8; ++w) M = fmaxf(M, red_m[w]);
      float S = 0.f;
      for (int w = 0; w < 8; ++w) S += (red_m[w] == -INFINITY) ? 0.f : red_s[w] * expf(red_m[w] - M);
      if (loss_row) {
        loss_row[q] = (M + logf(S)) - zy[q];
      } else {
        part_max[q] = M;
        part_sum[q] = S;
        for (int p = 1; p < n_split; ++p) {
          part_max[(long long)p * Q + q] = -INFINITY;
          part_sum[(long long)p * Q + q] = 0.f;
        }
      }
      if (repaired) atomicAdd(repaired, 1);
    }
    }
  }
}

// ---- top-k merge, sort network: one CTA per row, bitonic sort of 64-bit keys in shared memory ---------------------------
// key = order-preserving image of the score in the high word, ~index in the low word: DESCENDING key order is
// (score desc, index asc) = tf.nn.top_k order (loss.py:120).  Empty slots (idx < 0) carry key 0 and sort last.
// n_pad = next power of two >= n_part*k; O(n log^2 n) compare-exchanges instead of the O(n^2) counting kernel below
// (9 shards x 100: 28 k vs 810 k per row), and the lists are read and written once, coalesced.
__global__ void __launch_bounds__(256)
topk_merge_sort_kernel(const float* __restrict__ pv, const int* __restrict__ pi, int n_part, int Q, int k, int n_pad,
                       float* __restrict__ ov, int* __restrict__ oi) {
  extern __shared__ unsigned long long keys[];
  const int n = n_part * k;
  const int q = blockIdx.x;
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
    unsigned long long key = 0ull;
    if (i < n) {
      const int p = i / k, s = i % k;
      const long long o = ((long long)p * Q + q) * k + s;
      const int ii = pi[o];
      if (ii >= 0) key = topk_key(pv[o], ii);
    }
    keys[i] = key;
  }
  __syncthreads();
  bitonic_sort_desc(keys, n_pad);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const unsigned long long key = i < n_pad ? keys[i] : 0ull;
    const bool empty = key == 0ull;
    ov[(long long)q * k + i] = empty ? -INFINITY : topk_key_val(key);
    oi[(long long)q * k + i] = empty ? -1 : topk_key_idx(key);
  }
}

// ---- top-k merge: one CTA per row, rank-by-counting over the n_part*k candidates ----------------
// order: score desc, index asc (tf.nn.top_k); empty slots (idx < 0) sort last.
__global__ void __launch_bounds__(128)
topk_merge_kernel(const float* __restrict__ pv, const int* __restrict__ pi, int n_part, int Q, int k,
                  float* __restrict__ ov, int* __restrict__ oi) {
  extern __shared__ float sm[];
  const int n = n_part * k;
  float* v = sm;
  int* id = reinterpret_cast<int*>(sm + n);
  const int q = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int p = i / k, s = i % k;
    const long long o = ((long long)p * Q + q) * k + s;
    const int ii = pi[o];
    id[i] = ii < 0 ? 0x7fffffff : ii;
    v[i] = ii < 0 ? -INFINITY : pv[o];
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {   // default fill (fewer than k candidates)
    ov[(long long)q * k + i] = -INFINITY;
    oi[(long long)q * k + i] = -1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float vi = v[i];
    const int ii = id[i];
    if (ii == 0x7fffffff) continue;
    int pos = 0;
    for (int j = 0; j < n; ++j) {
      const float vj = v[j];
      const int ij = id[j];
      pos += (vj > vi || (vj == vi && ij < ii)) ? 1 : 0;
    }
    if (pos < k) {
      ov[(long long)q * k + pos] = vi;
      oi[(long long)q * k + pos] = ii;
    }
  }
}

// ---- masked two-level means (fixed summation order -> deterministic) -----------------------------
// Phase 1: one warp per user, lanes stride over the T positions (coalesced), fixed-shape shuffle tree ->
// user_part[b][8] = {loss, r@1, r@5, r@10, rr, rank_float}/(n_b + 1e-6), [n_b > 0], n_b.
// Phase 2: one CTA sums the users in a fixed order.
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
loss_metrics_user_kernel(const float* __restrict__ loss_row, const float* __restrict__ rank_row,
                         const int* __restrict__ row_of, const int* __restrict__ y_id, int B, int T, int item_num,
                         float* __restrict__ loss_bt, float* __restrict__ ranks, float* __restrict__ ranks_float,
                         float* __restrict__ user_part) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float fN = (float)item_num;
  float s[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // loss, r@1, r@5, r@10, rr, rank_float, n
  for (int t = lane; t < T; t += 32) {
    const long long i = (long long)b * T + t;
    const int row = row_of ? row_of[i] : (int)i;
    const bool m = y_id[i] > 0 && row >= 0;              // mask_y = sign(y_id)  (model.py:62)
    float l = 0.f, rk = 0.f, rf = 0.f;
    if (m) {
      l = loss_row ? loss_row[row] : 0.f;
      rk = rank_row ? rank_row[row] : 0.f;
      rf = rk / fN;                                       // loss.py:190
      s[0] += l;
      s[1] += (rk <= 0.f) ? 1.f : 0.f;                    // loss.py:194-196
      s[2] += (rk <= 4.f) ? 1.f : 0.f;
      s[3] += (rk <= 9.f) ? 1.f : 0.f;
      s[4] += 1.0f / (1.0f + rk);                         // loss.py:191
      s[5] += rf;
      s[6] += 1.f;
    }
    if (loss_bt) loss_bt[i] = l;
    if (ranks) ranks[i] = rk;
    if (ranks_float) ranks_float[i] = rf;
  }
#pragma unroll
  for (int i = 0; i < 7; ++i) s[i] = warp_sum_f(s[i]);
  if (lane == 0) {
    const float act = s[6] + 1e-6f;                       // model.py:114
    float* o = user_part + (long long)b * 8;
#pragma unroll
    for (int i = 0; i < 6; ++i) o[i] = s[i] / act;        // model.py:116, loss.py:208-213
    o[6] = (s[6] > 0.f) ? 1.f : 0.f;                      // user_count (model.py:113)
    o[7] = s[6];
  }
}

__global__ void __launch_bounds__(1024)
loss_metrics_final_kernel(const float* __restrict__ user_part, int B, float* __restrict__ scalars) {
  __shared__ float red[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float part[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) part[i] = 0.f;
  for (int b = threadIdx.x; b < B; b += 1024) {
    const float4 a = reinterpret_cast<const float4*>(user_part)[(long long)b * 2];
    const float4 c = reinterpret_cast<const float4*>(user_part)[(long long)b * 2 + 1];
    part[0] += a.x; part[1] += a.y; part[2] += a.z; part[3] += a.w;
    part[4] += c.x; part[5] += c.y; part[6] += c.z; part[7] += c.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[i] = warp_sum_f(part[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[i][warp] = part[i];
  }
  __syncthreads();
  if (warp == 0) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = warp_sum_f(red[i][lane]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) scalars[i] = v[i] / v[6];   // model.py:117, loss.py:215-219
      scalars[6] = v[6];
      scalars[7] = v[7];
    }
  }
}

}  // namespace htcn

using namespace htcn;

extern "C" int32_t htcn_prepare_wout(const float* w_out, const float* b_out, int32_t N, void* w_out_t, int32_t dtype,
                                     void* stream) {
  HTCN_REQUIRE(w_out && w_out_t && N > 0, "prepare_wout: bad args");
  dim3 block(32, 8);
  if (dtype == HTCN_BF16) {
    HTCN_REQUIRE(b_out, "prepare_wout: the bf16 layout folds the bias in, b_out is required");
    dim3 grid(ceil_div(N, 32), kDim / 32 + 1);
    prepare_wout_kernel<true><<<grid, block, 0, as_stream(stream)>>>(w_out, b_out, N, w_out_t);
  } else if (is_f32(dtype)) {
    const int pitch = planes_of(dtype) * kDim;
    dim3 grid(ceil_div(N, 32), pitch / 32);
    prepare_wout_kernel<false><<<grid, block, 0, as_stream(stream)>>>(w_out, b_out, N, w_out_t, pitch);
  } else {
    HTCN_REQUIRE(false, "prepare_wout: dtype %d", dtype);
  }
  HTCN_LAUNCH_CHECK("prepare_wout");
  return HTCN_OK;
}

extern "C" int32_t htcn_score_logits(const void* hout, int32_t hout_dtype, int32_t Q, const void* w_out_t,
                                     int32_t w_dtype, const float* b_out, int32_t n_items, float* logits,
                                     void* stream) {
  HTCN_REQUIRE(hout && w_out_t && logits && Q > 0 && n_items > 0, "score_logits: bad args");
  HTCN_REQUIRE(b_out || w_dtype == HTCN_BF16, "score_logits: b_out is NULL");
  const long long n = (long long)Q * n_items;
  HTCN_REQUIRE(hout_dtype != HTCN_F32_W256 || w_dtype != HTCN_BF16, "score_logits: 256-wide embeddings are fp32 only");
  score_logits_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(hout, hout_dtype == HTCN_BF16, Q, w_out_t,
                                                                      w_dtype == HTCN_BF16, b_out, n_items, logits,
                                                                      planes_of(hout_dtype));
  HTCN_LAUNCH_CHECK("score_logits");
  return HTCN_OK;
}

extern "C" int32_t htcn_target_logit(const void* hout, int32_t precision, int32_t Q, const void* w_out_t,
                                     const float* b_out, int32_t n_items, int32_t n0, const int32_t* y_id,
                                     float* target_logit, void* stream) {
  HTCN_REQUIRE(hout && w_out_t && y_id && target_logit && Q > 0 && n_items > 0, "target_logit: bad args");
  HTCN_REQUIRE(b_out || precision == HTCN_BF16, "target_logit: b_out is NULL");
  if (is_f32(precision))
    return target_logit_f32((const float*)hout, (const float*)w_out_t, b_out, y_id, Q, n_items, n0, target_logit,
                            as_stream(stream), planes_of(precision));
  if (precision == HTCN_BF16)
    return target_logit_bf16(hout, w_out_t, b_out, y_id, Q, n_items, n0, target_logit, as_stream(stream));
  HTCN_REQUIRE(false, "target_logit: precision %d", precision);
}

extern "C" int32_t htcn_score_ce_rank_topk(const void* hout, int32_t precision, int32_t Q, const void* w_out_t,
                                           const float* b_out, int32_t n_items, int32_t n0, const int32_t* y_id,
                                           float* target_logit, int32_t have_target, uint32_t flags, int32_t k,
                                           int32_t n_split, float* part_max, float* part_sum, int32_t* part_cnt,
                                           float* topk_val, int32_t* topk_idx, void* stream) {
  HTCN_REQUIRE(hout && w_out_t && Q > 0 && n_items > 0 && n_split >= 1, "score: bad args");
  HTCN_REQUIRE(b_out || precision == HTCN_BF16, "score: b_out is NULL");
  HTCN_REQUIRE(flags != 0 && (flags & ~7u) == 0, "score: flags 0x%x", flags);
  const bool need_t = flags & (HTCN_SCORE_CE | HTCN_SCORE_RANK);
  HTCN_REQUIRE(!need_t || (y_id && target_logit), "score: CE/RANK need y_id and target_logit");
  HTCN_REQUIRE(!(flags & HTCN_SCORE_CE) || (part_max && part_sum), "score: CE partial buffers NULL");
  HTCN_REQUIRE(!(flags & HTCN_SCORE_RANK) || part_cnt, "score: RANK partial buffer NULL");
  HTCN_REQUIRE(!(flags & HTCN_SCORE_TOPK) || (topk_val && topk_idx && k >= 1 && k <= HTCN_MAX_TOPK),
               "score: TOPK buffers NULL or k=%d out of [1,%d]", k, HTCN_MAX_TOPK);
  const int n_tiles = (n_items + 63) / 64;
  HTCN_REQUIRE(n_split <= n_tiles, "score: n_split=%d exceeds the number of 64-item tiles (%d)", n_split, n_tiles);
  cudaStream_t st = as_stream(stream);
  if (need_t && !have_target) {
    int32_t rc = htcn_target_logit(hout, precision, Q, w_out_t, b_out, n_items, n0, y_id, target_logit, stream);
    if (rc) return rc;
  }
  ScoreArgs a{hout, w_out_t, b_out, y_id, target_logit, part_max, part_sum, part_cnt, topk_val, topk_idx,
              Q, n_items, n0, k, n_split, flags};
  if (is_f32(precision)) {
    a.planes = planes_of(precision);
    return score_f32(a, st);
  }
  if (precision == HTCN_BF16) return score_bf16(a, st);
  HTCN_REQUIRE(false, "score: precision %d", precision);
}

// CE (+ rank) sweep of the bf16 tier with the exponent folded into the tensor-core product (k4_score_bf16.cu, kFoldFlag)
extern "C" int32_t htcn_score_ce_rank_folded(const void* hout, int32_t Q, const void* w_out_t, int32_t n_items, int32_t n0,
                                             const int32_t* y_id, const float* target_logit, uint32_t flags,
                                             int32_t n_split, float* part_max, float* part_sum, int32_t* part_cnt,
                                             void* workspace, void* stream) {
  HTCN_REQUIRE(hout && w_out_t && y_id && target_logit && part_max && part_sum && workspace && Q > 0 && n_items > 0 &&
                   n_split >= 1,
               "score_folded: bad args");
  HTCN_REQUIRE(flags == HTCN_SCORE_CE || flags == (HTCN_SCORE_CE | HTCN_SCORE_RANK), "score_folded: flags 0x%x", flags);
  HTCN_REQUIRE(!(flags & HTCN_SCORE_RANK) || part_cnt, "score_folded: RANK partial buffer NULL");
  const int n_tiles = (n_items + 63) / 64;
  HTCN_REQUIRE(n_split <= n_tiles, "score_folded: n_split=%d exceeds the number of 64-item tiles (%d)", n_split, n_tiles);
  ScoreArgs a{hout, w_out_t, nullptr, y_id, target_logit, part_max, part_sum, part_cnt, nullptr, nullptr,
              Q, n_items, n0, 0, n_split, flags};
  return score_ce_rank_folded_bf16(a, workspace, as_stream(stream));
}

extern "C" int64_t htcn_topk_workspace_bytes(int32_t precision, int32_t Q, int32_t n_items, int32_t k, int32_t n_split) {
  if (topk_two_pass_ok(precision, n_items, k)) return topk_workspace_bytes_bf16(Q, n_items, k, n_split);
  return (int64_t)topk_heap_splits(precision, n_items, n_split) * Q * k * 8 + 256;
}

extern "C" int32_t htcn_score_topk(const void* hout, int32_t precision, int32_t Q, const void* w_out_t, const float* b_out,
                                   int32_t n_items, int32_t n0, int32_t k, int32_t n_split, void* workspace,
                                   int64_t workspace_bytes, float* out_val, int32_t* out_idx, int32_t* overflow_rows,
                                   void* stream) {
  HTCN_REQUIRE(hout && w_out_t && workspace && out_val && out_idx && Q > 0 && n_items > 0, "score_topk: bad args");
  HTCN_REQUIRE(k >= 1 && k <= HTCN_MAX_TOPK, "score_topk: k=%d out of [1,%d]", k, HTCN_MAX_TOPK);
  HTCN_REQUIRE(is_f32(precision) || precision == HTCN_BF16, "score_topk: precision %d", precision);
  HTCN_REQUIRE(workspace_bytes >= htcn_topk_workspace_bytes(precision, Q, n_items, k, n_split),
               "score_topk: workspace too small (%lld bytes)", (long long)workspace_bytes);
  cudaStream_t st = as_stream(stream);
  if (topk_two_pass_ok(precision, n_items, k))
    return score_topk_bf16(hout, Q, w_out_t, n_items, n0, k, n_split, workspace, workspace_bytes, out_val, out_idx,
                           overflow_rows, st);
  // heap sweep into [ns, Q, k] partial lists, then merge
  const int ns = topk_heap_splits(precision, n_items, n_split);
  float* tv = reinterpret_cast<float*>(workspace);
  int32_t* ti = reinterpret_cast<int32_t*>(tv + (size_t)ns * Q * k);
  if (overflow_rows) HTCN_CUDA(cudaMemsetAsync(overflow_rows, 0, 4, st));
  int32_t rc = htcn_score_ce_rank_topk(hout, precision, Q, w_out_t, b_out, n_items, n0, nullptr, nullptr, 1,
                                       HTCN_SCORE_TOPK, k, ns, nullptr, nullptr, nullptr, tv, ti, stream);
  if (rc) return rc;
  return htcn_topk_merge(tv, ti, ns, Q, k, out_val, out_idx, stream);
}

extern "C" int32_t htcn_score_ce_rank_topk_fused(const void* hout, int32_t precision, int32_t Q, const void* w_out_t,
                                                 const float* b_out, int32_t n_items, int32_t n0, const int32_t* y_id,
                                                 const float* target_logit, int32_t k, int32_t n_split, void* workspace,
                                                 int64_t workspace_bytes, float* part_max, float* part_sum,
                                                 int32_t* part_cnt, float* out_val, int32_t* out_idx,
                                                 int32_t* overflow_rows, void* stream) {
  HTCN_REQUIRE(hout && w_out_t && y_id && target_logit && part_max && part_sum && part_cnt && out_val && out_idx && workspace,
               "score_fused: NULL pointer");
  HTCN_REQUIRE(Q > 0 && n_items > 0 && n_split >= 1 && k >= 1 && k <= HTCN_MAX_TOPK, "score_fused: Q=%d n_items=%d k=%d", Q,
               n_items, k);
  if (topk_two_pass_ok(precision, n_items, k)) {
    HTCN_REQUIRE(workspace_bytes >= topk_workspace_bytes_bf16(Q, n_items, k, n_split), "score_fused: workspace too small");
    ScoreArgs a{hout, w_out_t, b_out, y_id, target_logit, part_max, part_sum, part_cnt, nullptr, nullptr,
                Q, n_items, n0, 0, n_split, HTCN_SCORE_CE | HTCN_SCORE_RANK};
    return score_topk_bf16(hout, Q, w_out_t, n_items, n0, k, n_split, workspace, workspace_bytes, out_val, out_idx,
                           overflow_rows, as_stream(stream), &a);
  }
  // fp32 tier / small shards: the loss sweep and the heap top-k as separate sweeps
  int32_t rc = htcn_score_ce_rank_topk(hout, precision, Q, w_out_t, b_out, n_items, n0, y_id, const_cast<float*>(target_logit),
                                       1, HTCN_SCORE_CE | HTCN_SCORE_RANK, 0, n_split, part_max, part_sum, part_cnt, nullptr,
                                       nullptr, stream);
  if (rc) return rc;
  return htcn_score_topk(hout, precision, Q, w_out_t, b_out, n_items, n0, k, n_split, workspace, workspace_bytes, out_val,
                         out_idx, overflow_rows, stream);
}

extern "C" int32_t htcn_score_finish(const float* part_max, const float* part_sum, const int32_t* part_cnt,
                                     int32_t n_part, int32_t Q, const int32_t* y_id, const float* target_logit,
                                     float* loss_row, float* rank_row, void* stream) {
  HTCN_REQUIRE(Q > 0 && n_part >= 1, "score_finish: bad sizes");
  HTCN_REQUIRE(!loss_row || (part_max && part_sum && target_logit), "score_finish: CE inputs NULL");
  HTCN_REQUIRE(!rank_row || part_cnt, "score_finish: rank input NULL");
  score_finish_kernel<<<ceil_div(Q, 256), 256, 0, as_stream(stream)>>>(part_max, part_sum, part_cnt, n_part, Q,
                                                                      y_id, target_logit, loss_row, rank_row);
  HTCN_LAUNCH_CHECK("score_finish");
  return HTCN_OK;
}

extern "C" int32_t htcn_score_ce_repair(const void* hout, int32_t precision, int32_t Q, const void* w_out_t,
                                        int32_t n_items, const float* target_logit, float* loss_row, int32_t* repaired,
                                        void* stream) {
  HTCN_REQUIRE(hout && w_out_t && target_logit && loss_row && Q > 0 && n_items > 0, "score_ce_repair: bad args");
  if (is_f32(precision)) return HTCN_OK;               // the fp32 sweep keeps a running max: nothing to repair
  HTCN_REQUIRE(precision == HTCN_BF16, "score_ce_repair: precision %d", precision);
  const int chunks = ceil_div(Q, 256);
  const int grid = chunks < 148 * 8 ? chunks : 148 * 8;
  ce_repair_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(hout), Q,
                                                            reinterpret_cast<const __nv_bfloat16*>(w_out_t), n_items,
                                                            target_logit, loss_row, nullptr, nullptr, 0, repaired);
  HTCN_LAUNCH_CHECK("ce_repair_bf16_kernel");
  return HTCN_OK;
}

extern "C" int32_t htcn_score_ce_repair_shard(const void* hout, int32_t precision, int32_t Q, const void* w_out_t,
                                              int32_t n_items, float* part_max, float* part_sum, int32_t n_split,
                                              int32_t* repaired, void* stream) {
  HTCN_REQUIRE(hout && w_out_t && part_max && part_sum && Q > 0 && n_items > 0 && n_split >= 1,
               "score_ce_repair_shard: bad args");
  if (is_f32(precision)) return HTCN_OK;               // the fp32 sweep keeps a running max per split
  HTCN_REQUIRE(precision == HTCN_BF16, "score_ce_repair_shard: precision %d", precision);
  const int chunks = ceil_div(Q, 256);
  const int grid = chunks < 148 * 8 ? chunks : 148 * 8;
  ce_repair_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(hout), Q,
                                                            reinterpret_cast<const __nv_bfloat16*>(w_out_t), n_items,
                                                            nullptr, nullptr, part_max, part_sum, n_split, repaired);
  HTCN_LAUNCH_CHECK("ce_repair_bf16_kernel(shard)");
  return HTCN_OK;
}

extern "C" int32_t htcn_topk_merge(const float* part_val, const int32_t* part_idx, int32_t n_part, int32_t Q,
                                   int32_t k, float* out_val, int32_t* out_idx, void* stream) {
  HTCN_REQUIRE(part_val && part_idx && out_val && out_idx && n_part >= 1 && Q > 0 && k >= 1, "topk_merge: bad args");
  {
    int n_pad = 2;
    while (n_pad < n_part * k) n_pad <<= 1;
    if (n_pad <= 16384) {                                  // 128 KB of keys: the sort network
      const size_t smem_sort = (size_t)n_pad * 8;
      HTCN_CUDA(cudaFuncSetAttribute(topk_merge_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sort));
      const int threads = n_pad / 2 < 256 ? (n_pad / 2 < 32 ? 32 : n_pad / 2) : 256;
      topk_merge_sort_kernel<<<Q, threads, smem_sort, as_stream(stream)>>>(part_val, part_idx, n_part, Q, k, n_pad, out_val,
                                                                        out_idx);
      HTCN_LAUNCH_CHECK("topk_merge_sort");
      return HTCN_OK;
    }
  }
  const size_t smem = (size_t)n_part * k * 8;
  HTCN_REQUIRE(smem <= 200 * 1024, "topk_merge: n_part*k=%d candidates exceed shared memory", n_part * k);
  HTCN_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_merge_kernel<<<Q, 128, smem, as_stream(stream)>>>(part_val, part_idx, n_part, Q, k, out_val, out_idx);
  HTCN_LAUNCH_CHECK("topk_merge");
  return HTCN_OK;
}

extern "C" int32_t htcn_loss_metrics_reduce(const float* loss_row, const float* rank_row, const int32_t* row_of,
                                            const int32_t* y_id, int32_t B, int32_t T, int32_t item_num,
                                            float* loss_bt, float* ranks, float* ranks_float, float* user_part,
                                            float* scalars, void* stream) {
  HTCN_REQUIRE(y_id && scalars && user_part && B > 0 && T > 0 && item_num > 0, "loss_metrics_reduce: bad args");
  loss_metrics_user_kernel<<<ceil_div(B, 8), 256, 0, as_stream(stream)>>>(loss_row, rank_row, row_of, y_id, B, T, item_num,
                                                                        loss_bt, ranks, ranks_float, user_part);
  HTCN_LAUNCH_CHECK("loss_metrics_user_kernel");
  loss_metrics_final_kernel<<<1, 1024, 0, as_stream(stream)>>>(user_part, B, scalars);
  HTCN_LAUNCH_CHECK("loss_metrics_final_kernel");
  return HTCN_OK;
}
