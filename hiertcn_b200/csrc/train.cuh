// Internal building blocks of the training step (backward + Adam), fp32 on the FMA pipe.
// The reference trains with tf.train.AdamOptimizer(lr).minimize(loss) (model.py:134-141); TensorFlow's autodiff
// produces dense gradients for every variable (the one-hot matmuls make even the embedding gradients dense).
#pragma once

#include "common.cuh"

namespace htcn {

// one level of the fp32 conv stack (k2_tcn_f32.cu), also used transposed / anti-causal by the backward
struct LevelArgs {
  const void* in;             // [R,128] f32 (or bf16 when in_bf16)
  const float* w;             // [K,128,128]
  const float* bias;          // [128] or NULL
  const float* sbias;         // [S,B,128] or NULL
  void* out;                  // [R or n_out,128]
  const int* out_row;         // [R] or NULL
  long long R;
  int T, B, K, dil;
  int conv_epilogue;          // 1: relu, + in[r,:] (identity residual), relu; 0: linear; 2: + resid[r,:] (backward: data
                              // gradient); 3: relu, + resid[r,:], relu (down-sample residual, customized_tcn_cell.py:102-106,123-124)
  int in_bf16, out_bf16;
  float* aux;                 // [R,128] relu(conv + bias) before the residual add, saved for the backward, or NULL
  const float* resid;         // [R,128] added in epilogues 2 and 3 (may alias out in 2)
  const float* drop;          // training dropout (customized_tcn_cell.py:100,119: noise_shape [1,1,C], one channel mask per
  int drop_stride;            //   slot and level): relu(conv + b) is multiplied by drop[slot(r) * drop_stride + c]; NULL = off
  int in_planes;              // 0/1: plain.  P > 1: the input is P 128-wide planes ([P][R][128], in_plane_stride floats apart),
  long long in_plane_stride;  //   run as P*K taps: tap = pi*K + kt reads plane pi shifted by (K-1-kt)*dil, w = [.., P*K, 128, 128]
  long long out_plane_stride; // grid.y = output plane po: w += po*P*K*128*128, bias += po*128, out / resid / identity residual plane po
  long long resid_plane_stride;
  int anti;                   // 1: taps read r + shift, zero beyond the END of r's sequence (transposed convolution)
  int w_nt;                   // 1: every W[tap] is applied transposed
  void* tc_ws;                // non-NULL: run the level on the tensor cores (k2_level_tc.cu) when its shape allows; device
                              //   workspace of HTCN_K2TC_WS_BYTES for the level's bf16 weight tiles
  int tc_split;               // tensor-core path: 1 = fp32-grade products (bf16 hi/lo split, 3 MMAs), 0 = plain bf16
};
constexpr long long HTCN_K2TC_WS_BYTES = 8LL * 2 * 128 * 128 * 2;      // K <= 8 taps x (hi, lo) x 32 KB
// dispatches to k2_level_tc_launch when a.tc_ws is set and k2_level_tc_supported(a), else the FFMA kernel
int32_t k2_level_launch(const LevelArgs& a, const SlotTable& slots, cudaStream_t st);
bool k2_level_tc_supported(const LevelArgs& a, const SlotTable& slots);
int32_t k2_level_tc_launch(const LevelArgs& a, const SlotTable& slots, cudaStream_t st);

// fused tcgen05 conv stack (k2_tcn_bf16.cu); h_save / a_save (bf16, optional) keep every layer's output and every
// level's pre-residual activation for the backward
// ds_w / ds_b: per level, the 1x1 down-sample residual Dense of customized_tcn_cell.py:102-106 (entry NULL = identity
// residual; the arrays themselves may be NULL)
int32_t tcn_forward_bf16(const void* xe, int xe_dtype, const float* w_in_x, const float* sbias,
                         const float* const* conv_w, const float* const* conv_b, const float* const* ds_w,
                         const float* const* ds_b, int n_levels, int K,
                         const SlotTable& slots, int B, int T, const int* out_row, void* hout, int hout_dtype,
                         float* scratch, cudaStream_t st, void* h_save = nullptr, void* a_save = nullptr,
                         const float* drop = nullptr /* [S][n_levels][128] dropout scales (training) */);

// C[M,N] (+)= A[M,K] * op(B).  trans_b = 0: B is [K,N] with leading dimension ldb; trans_b = 1: B is [N,K].
// N % 128 == 0, K % 16 == 0, all leading dimensions multiples of 4 floats, 16-byte aligned bases.
int32_t sgemm(bool trans_b, long long M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
              int ldc, bool accumulate, cudaStream_t st);

// Split-K "TN" product with atomic accumulation, the weight-gradient form:
//     C[c, f] += sum_{r < R} A[src(r), c] * D[r, f]        c, f in [0,128)
// src(r) = r - shift, taken as a zero row when r's position inside its sequence is < shift (the causal left pad of
// customized_tcn_cell.py:46-49); shift = 0 (and slots == nullptr) for plain products.
// A is fp32, or bf16 when a_bf16 (then lda counts bf16 elements).
int32_t sgemm_tn_atomic(long long R, const void* A, int lda, const float* D, int ldd, float* C, int ldc, int shift,
                        int T, const SlotTable* slots, cudaStream_t st, bool a_bf16 = false);

// ---- tensor-core weight gradients (bwd_wgrad_bf16.cu) ----
// zero-padded, transposed bf16 copy of a [B*T,128] activation: [128][Kp], every sequence preceded by P zero columns
struct PadGeom {
  int n_slots, B, T, P;
  int off[HTCN_MAX_SLOTS + 1];
  long long base[HTCN_MAX_SLOTS + 1];    // first column of slot s
  long long Kp;                          // columns, multiple of 64
};
PadGeom make_pad_geom(const SlotTable& slots, int B, int T, int P);
// n_shifts copies, copy i shifted by shifts[i] columns inside every sequence (zero before its start), dst_stride elements apart
int32_t pad_transpose_bf16(const void* src, bool src_bf16, const PadGeom& g, const int* shifts, int n_shifts, void* dst,
                           long long dst_stride, cudaStream_t st);
// dW[t][cin][cout] += sum_col aT_t[cin][col] * bT[cout][col]; aT_t = aT + t * a_stride (tap t's pre-shifted copy), bT: bf16
// [128][Kp]; dW fp32, atomics
int32_t wgrad_bf16(const void* aT, long long a_stride, const void* bT, long long Kp, int n_taps, float* dW, cudaStream_t st);

// the same with MN-major operands (no transposes): row-major zero-padded bf16 copies [Kp][128]; the tap shift is a TMA row
// coordinate, one copy serves every tap:  dW[t][cin][cout] += sum_r aP[r - shifts[t]][cin] * bP[r][cout]
int32_t pad_rows_bf16(const void* src, bool src_bf16, const PadGeom& g, void* dst, cudaStream_t st);
int32_t wgrad_mn_bf16(const void* aP, const void* bP, long long Kp, const int* shifts, int n_taps, float* dW, cudaStream_t st);

// out[f] += sum_r D[r, f]  (f < cols, cols <= 256)
int32_t colsum_atomic(long long R, const float* D, int ldd, int cols, float* out, cudaStream_t st);

}  // namespace htcn
