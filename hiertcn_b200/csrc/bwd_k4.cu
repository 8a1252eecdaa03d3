// K4 backward (fp32): gradient of the full-catalog softmax cross-entropy (loss.py:20-21 under the masks and means of
// model.py:105-117) with respect to the user embeddings Hout, the output table W_out^T and its bias, WITHOUT ever
// materialising the [Q, N] logits or their gradient.  For a tile of 128 scored rows and a tile of 128 items:
//
//     Z  = H Wt^T + b                                  (recomputed, 128 x 128)
//     P  = g_row * (exp(Z - lse_row) - [j == y_row])   (dL/dZ; lse from the forward sweep)
//     dH  += P  Wt                                     (kept in registers across the item loop)
//     dWt += P^T H,   db += colsum(P)                  (atomic accumulation across row tiles)
//
// The three products run on the FMA pipe with 8x8 register tiles out of natural-layout shared-memory tiles whose
// rows are padded to 132 floats (conflict-free float4 reads along either index).  What TensorFlow does instead: the
// dense [B,T,N] softmax gradient tensor and two cuBLAS-sized GEMMs on it.
#include "train.cuh"

namespace htcn {

namespace {
constexpr int kPitch = 132;
constexpr int kTile = 128;
constexpr int kThreads = 256;
constexpr size_t kSmemBytes = sizeof(float) * (3 * kTile * kPitch + 3 * kTile) + sizeof(int) * kTile;

struct K4BwdArgs {
  const void* hout; int hout_bf16;
  const float* wt; const float* b_out;
  const int* y_id; const float* loss_row; const float* zy; const float* g_row;
  float* d_hout; float* d_wt; float* d_b;
  int Q, n_items, n0, n_split;
};

__global__ void __launch_bounds__(kThreads, 1) k4_ce_backward_f32(K4BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Hs = smem;                          // [128][132]  rows of Hout
  float* Ws = Hs + kTile * kPitch;           // [128][132]  rows of W_out^T (items)
  float* Ps = Ws + kTile * kPitch;           // [128][132]  dL/dZ tile, [row][item]
  float* lse_s = Ps + kTile * kPitch;        // [128]
  float* g_s = lse_s + kTile;                // [128]
  float* b_s = g_s + kTile;                  // [128]
  int* y_s = reinterpret_cast<int*>(b_s + kTile);

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long r0 = (long long)blockIdx.x * kTile;
  // this CTA's slice of the catalog: whole item tiles, split evenly
  const int tiles = (a.n_items + kTile - 1) / kTile;
  const int per = (tiles + a.n_split - 1) / a.n_split;
  const int t_begin = blockIdx.y * per;
  const int t_end = min(tiles, t_begin + per);
  if (t_begin >= t_end) return;

  for (int e = tid; e < kTile * 32; e += kThreads) {       // 128 rows x 32 float4
    const int r = e >> 5, c = (e & 31) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < a.Q) {
      if (a.hout_bf16) {
        const uint2 q = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(a.hout) + (r0 + r) * kDim + c);
        v = make_float4(bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y));
      } else {
        v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.hout) + (r0 + r) * kDim + c);
      }
    }
    *reinterpret_cast<float4*>(Hs + r * kPitch + c) = v;
  }
  if (tid < kTile) {
    const bool ok = r0 + tid < a.Q;
    lse_s[tid] = ok ? a.loss_row[r0 + tid] + a.zy[r0 + tid] : 0.f;
    g_s[tid] = ok ? a.g_row[r0 + tid] : 0.f;
    y_s[tid] = ok ? a.y_id[r0 + tid] : -1;
  }

  float dh[8][8];      // rows {ty + 16 i}, dims {tx*4.., 64 + tx*4..}
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dh[i][j] = 0.f;

  for (int t = t_begin; t < t_end; ++t) {
    const int j0 = t * kTile;
    __syncthreads();                                        // previous tile fully consumed (and Hs visible)
    for (int e = tid; e < kTile * 32; e += kThreads) {
      const int r = e >> 5, c = (e & 31) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + r < a.n_items) v = __ldg(reinterpret_cast<const float4*>(a.wt + (long long)(j0 + r) * kDim + c));
      *reinterpret_cast<float4*>(Ws + r * kPitch + c) = v;
    }
    if (tid < kTile) b_s[tid] = (j0 + tid < a.n_items) ? __ldg(a.b_out + j0 + tid) : 0.f;
    __syncthreads();

    // ---- Z tile: rows {ty + 16 i} x items {tx + 16 j} ------------------------------------------------------
    {
      float z[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) z[i][j] = 0.f;
#pragma unroll 2
      for (int k = 0; k < kDim; k += 4) {
        float4 av[8], bv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = *reinterpret_cast<const float4*>(Hs + (ty + 16 * i) * kPitch + k);
#pragma unroll
        for (int j = 0; j < 8; ++j) bv[j] = *reinterpret_cast<const float4*>(Ws + (tx + 16 * j) * kPitch + k);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            z[i][j] = fmaf(av[i].x, bv[j].x, z[i][j]);
            z[i][j] = fmaf(av[i].y, bv[j].y, z[i][j]);
            z[i][j] = fmaf(av[i].z, bv[j].z, z[i][j]);
            z[i][j] = fmaf(av[i].w, bv[j].w, z[i][j]);
          }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = ty + 16 * i;
        const float lse = lse_s[r], g = g_s[r];
        const int y = y_s[r];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int jl = tx + 16 * j;
          float p = 0.f;
          if (j0 + jl < a.n_items) p = g * (expf(z[i][j] + b_s[jl] - lse) - ((a.n0 + j0 + jl == y) ? 1.f : 0.f));
          Ps[r * kPitch + jl] = p;
        }
      }
    }
    __syncthreads();

    // ---- dH += P Wt : rows {ty + 16 i} x dims {tx*4.., 64 + tx*4..} ---------------------------------------
#pragma unroll 1
    for (int j = 0; j < kTile; j += 4) {
      float4 pv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pv[i] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * i) * kPitch + j);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 w0 = *reinterpret_cast<const float4*>(Ws + (j + jj) * kPitch + tx * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(Ws + (j + jj) * kPitch + 64 + tx * 4);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float p = (jj == 0) ? pv[i].x : (jj == 1) ? pv[i].y : (jj == 2) ? pv[i].z : pv[i].w;
#pragma unroll
          for (int d = 0; d < 8; ++d) dh[i][d] = fmaf(p, wv[d], dh[i][d]);
        }
      }
    }

    // ---- dWt tile = P^T H : items {ty*4.., 64 + ty*4..} x dims {tx*4.., 64 + tx*4..}; db = colsum(P) -------
    {
      float dw[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dw[i][j] = 0.f;
#pragma unroll 4
      for (int r = 0; r < kTile; ++r) {
        const float4 p0 = *reinterpret_cast<const float4*>(Ps + r * kPitch + ty * 4);
        const float4 p1 = *reinterpret_cast<const float4*>(Ps + r * kPitch + 64 + ty * 4);
        const float4 h0 = *reinterpret_cast<const float4*>(Hs + r * kPitch + tx * 4);
        const float4 h1 = *reinterpret_cast<const float4*>(Hs + r * kPitch + 64 + tx * 4);
        const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int d = 0; d < 8; ++d) dw[i][d] = fmaf(pv[i], hv[d], dw[i][d]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int jl = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4);
        if (j0 + jl >= a.n_items) continue;
#pragma unroll
        for (int dhf = 0; dhf < 2; ++dhf)
          atomicAdd(reinterpret_cast<float4*>(a.d_wt + (long long)(j0 + jl) * kDim + dhf * 64 + tx * 4),
                    make_float4(dw[i][dhf * 4 + 0], dw[i][dhf * 4 + 1], dw[i][dhf * 4 + 2], dw[i][dhf * 4 + 3]));
      }
      if (a.d_b && tid < kTile && j0 + tid < a.n_items) {
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < kTile; ++r) s += Ps[r * kPitch + tid];
        atomicAdd(a.d_b + j0 + tid, s);
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = r0 + ty + 16 * i;
    if (r >= a.Q) continue;
#pragma unroll
    for (int dhf = 0; dhf < 2; ++dhf)
      atomicAdd(reinterpret_cast<float4*>(a.d_hout + r * kDim + dhf * 64 + tx * 4),
                make_float4(dh[i][dhf * 4 + 0], dh[i][dhf * 4 + 1], dh[i][dhf * 4 + 2], dh[i][dhf * 4 + 3]));
  }
}

}  // namespace
}  // namespace htcn

extern "C" int32_t htcn_score_ce_backward(const void* hout, int32_t hout_dtype, int32_t Q, const float* wt,
                                          const float* b_out, int32_t n_items, int32_t n0, const int32_t* y_id,
                                          const float* loss_row, const float* target_logit, const float* g_row,
                                          float* d_hout, float* d_wt, float* d_b, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(hout && wt && b_out && y_id && loss_row && target_logit && g_row && d_hout && d_wt,
               "score_ce_backward: NULL pointer");
  HTCN_REQUIRE(hout_dtype == HTCN_F32 || hout_dtype == HTCN_BF16, "score_ce_backward: hout dtype %d", hout_dtype);
  HTCN_REQUIRE(Q >= 0 && n_items > 0, "score_ce_backward: Q=%d n_items=%d", Q, n_items);
  if (Q == 0) return HTCN_OK;
  cudaStream_t st = as_stream(stream);
  HTCN_CUDA(cudaMemsetAsync(d_hout, 0, sizeof(float) * (size_t)Q * kDim, st));
  K4BwdArgs a{hout, hout_dtype == HTCN_BF16, wt, b_out, y_id, loss_row, target_logit, g_row, d_hout, d_wt, d_b,
              Q, n_items, n0, 1};
  const int row_tiles = ceil_div(Q, kTile), item_tiles = ceil_div(n_items, kTile);
  int ns = ceil_div(2 * 148, row_tiles);                    // at least two waves of CTAs
  if (ns > item_tiles) ns = item_tiles;
  if (ns < 1) ns = 1;
  a.n_split = ns;
  HTCN_CUDA(cudaFuncSetAttribute(k4_ce_backward_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  k4_ce_backward_f32<<<dim3(row_tiles, ns), kThreads, kSmemBytes, st>>>(a);
  HTCN_LAUNCH_CHECK("k4_ce_backward_f32");
  return HTCN_OK;
}
