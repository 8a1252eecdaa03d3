// K4 backward on the tensor cores (bf16 operands, fp32 accumulation): the gradient of the full-catalog softmax
// cross-entropy without materialising the [Q, N] logits or their gradient, as two passes of ONE kernel shaped like a
// flash-attention forward:
//
//     S = X Y^T        (tcgen05, 2 x (128 x 64) per iteration, K = 128, accumulators in TMEM, double-buffered)
//     P = f(S)         (epilogue warps: TMEM -> registers -> exp / one-hot / scale -> bf16 pairs -> TMEM, over S)
//     O += P V^T       (tcgen05 with A from TMEM, 2 x (128 x 128), K = 64; O stays in TMEM for the whole sweep)
//
//   pass A (dHout): X = Hout rows (256 per CTA, stationary), Y = W_out^T rows (items, streamed), V = W_out [128, N];
//                   P[q, j] = exp(z - lse_q) - [j == y_q];  dHout[q] = g_q * O[q]
//   pass B (dW^T):  X = W_out^T rows (256 items per CTA, stationary), Y = Hout rows (queries, streamed),
//                   V = Hout^T [128, Q];  P[j, q] = g_q (exp(z - lse_q) - [j == y_q]);  dW^T[j] += O[j];
//                   db[j] += sum_q P[j, q]
//
// with z = S + b and lse_q = loss_q + z_{q, y_q} from the forward sweep.  "TMEM lane = stationary row": every epilogue
// thread owns one row of S, so P goes back to tensor memory as the A operand of the second product with one tcgen05.st
// per thread, and db is a thread-local sum.  Replaces TensorFlow's dense [B,T,N] softmax gradient + two GEMMs on it
// (model_tcn.py:41, loss.py:20-21, model.py:134-141).
#include "common.cuh"
#include "sm100.cuh"

#include <cstddef>

namespace htcn {
using namespace sm100;

namespace {
constexpr int kBM = 128;          // TMEM lanes = rows of one stationary half
constexpr int kHalves = 2;        // a CTA keeps 2 x 128 stationary rows, so every streamed tile (32 KB out of L2) is used
                                  // twice: at one half per CTA the kernel ran into the L2 bandwidth ceiling (~43 B/clk/SM)
constexpr int kBN = 64;           // streamed rows per iteration
constexpr int kStages = 4;
constexpr int kEpiWarps = 16;     // (half, 32-column group) x 4 TMEM lane quarters
constexpr int kColsPerWarp = 32;
constexpr int kThreads = 64 + 32 * kEpiWarps;
// TMEM map (512 columns): S[half][buf] = 64 fp32 columns at (half*2 + buf)*64; O[half] = 128 columns at 256 + half*128.
// P is a TMEM-resident A operand of the second product (tcgen05.mma with A from tensor memory, two bf16 per 32-bit
// column) and ALIASES S: the epilogue warp that has read columns [32 cg, 32 cg + 32) of S(half, buf) writes its 16 packed
// columns back to [32 cg, 32 cg + 16).  The tensor pipe executes in issue order, so S(i+2) -- issued after the product that
// consumes P(i) -- cannot overwrite P(i) early, and no "accumulator free" barriers are needed.
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemO = 256;
constexpr float kLog2e = 1.4426950408889634f;

struct alignas(1024) BwdSmem {
  uint8_t x[kHalves][2][kBM * 128];    // stationary operand: [half][K chunk 0..63 / 64..127]
  uint8_t y[kStages][2][kBN * 128];    // streamed operand of S = X Y^T
  uint8_t v[kStages][kBM * 128];       // streamed operand of O += P V^T: [128 dims][64 streamed indices]
  float colA[kStages][kBN];            // per streamed index, bulk-copied with the tiles (BwdArgs::bL / lseL / gq / yq)
  float colB[kStages][kBN];
  int colI[kStages][kBN];
  uint64_t x_full, full[kStages], empty[kStages], s_full[2], p_full[2], o_full;
  uint32_t tmem_base;
};

struct BwdArgs {
  // per-index vectors prepared by k4_bwd_prep_kernel, padded to whole 64-index tiles:
  //   bL [ceil64(n_items)] = b_out * log2e (-inf padding);  lseL [ceil64(Q)] = (loss_row + zy) * log2e (+inf padding);
  //   gq [ceil64(Q)] = g_row (0 padding);  yq [ceil64(Q)] = y_id (-2 padding)
  const float* bL;
  const float* lseL;
  const float* gq;
  const int* yq;
  float* out;               // pass A: d_hout [Q,128]; pass B: d_wt [n_items,128]   (atomic accumulation)
  float* d_b;               // pass B: [n_items] or NULL
  float* lsum;              // pass F: [Q] sum_j exp(z_j - z_y), atomic accumulation (lseL then holds z_y * log2e)
  int Q, n_items, n0;
  int n_stream;             // rows of the streamed operand (pass A: n_items, pass B: Q)
  int n_split;
};

// kMode 0: pass A, 1: pass B, 2: pass F = forward + pass A in one sweep.  With the TARGET logit as the reference point
// (the forward sweep's convention: loss = log sum_j exp(z_j - z_y), k4_score_bf16.cu) the softmax numerator needs no
// running maximum, so O = sum_j exp(z_j - z_y) w_j and l = sum_j exp(z_j - z_y) accumulate without rescaling and
//     loss_q = log l_q,     dHout[q] = g_q (O_q / l_q - w_{y_q})
// come out of ONE catalog sweep (ce_fused_finish_kernel): the separate forward sweep that only produced lse is gone.
template <int kMode>
__global__ void __launch_bounds__(kThreads, 1)
k4_ce_backward_bf16(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y,
                    const __grid_constant__ CUtensorMap tmap_v, BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  auto& sm = *reinterpret_cast<BwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * (kHalves * kBM);                       // first stationary row
  const int tiles_all = (a.n_stream + kBN - 1) / kBN;
  const int t_begin = (int)((long long)tiles_all * blockIdx.y / a.n_split);
  const int t_end = (int)((long long)tiles_all * (blockIdx.y + 1) / a.n_split);
  const int n_iter = t_end - t_begin;
  if (n_iter <= 0) return;
  constexpr bool kPassB = kMode == 1;
  constexpr bool kPassF = kMode == 2;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_y);
    prefetch_tmap(&tmap_v);
    mbar_init(&sm.x_full, 1);
    mbar_init(&sm.o_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm.s_full[s], 1);
      mbar_init(&sm.p_full[s], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(&sm.tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, sm.tmem_base, 0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(&sm.x_full, kHalves * 2 * kBM * 128);
      for (int h = 0; h < kHalves; ++h) {
        tma_load_2d(sm.x[h][0], &tmap_x, 0, m0 + h * kBM, &sm.x_full);
        tma_load_2d(sm.x[h][1], &tmap_x, 64, m0 + h * kBM, &sm.x_full);
      }
      for (int i = 0; i < n_iter; ++i) {
        const int s = i % kStages;
        mbar_wait(&sm.empty[s], ((i / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&sm.full[s], 2 * kBN * 128 + kBM * 128 + (kPassB ? 3 : 1) * kBN * 4);
        const int j0 = (t_begin + i) * kBN;                          // rows / columns beyond the tensor are zero-filled
        tma_load_2d(sm.y[s][0], &tmap_y, 0, j0, &sm.full[s]);
        tma_load_2d(sm.y[s][1], &tmap_y, 64, j0, &sm.full[s]);
        tma_load_2d(sm.v[s], &tmap_v, j0, 0, &sm.full[s]);
        if (!kPassB) {
          bulk_load_1d(sm.colA[s], a.bL + j0, kBN * 4, &sm.full[s]);
        } else {
          bulk_load_1d(sm.colA[s], a.lseL + j0, kBN * 4, &sm.full[s]);
          bulk_load_1d(sm.colB[s], a.gq + j0, kBN * 4, &sm.full[s]);
          bulk_load_1d(sm.colI[s], a.yq + j0, kBN * 4, &sm.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this (warp-uniform waits and descriptor arithmetic); one elected lane issues.
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc_s = make_idesc_bf16(kBM, kBN);
      constexpr uint32_t idesc_o = make_idesc_bf16(kBM, kBM);
      const uint32_t x_addr = smem_u32(sm.x[0][0]), y_addr = smem_u32(sm.y[0][0]), v_addr = smem_u32(sm.v[0]);
      auto issue_s = [&](int i) {                                    // S[h][i & 1] = X_h Y_i^T for both halves
        const int s = i % kStages, buf = i & 1;
        mbar_wait(&sm.full[s], (i / kStages) & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int h = 0; h < kHalves; ++h)
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t da = make_desc_k_sw128(x_addr + (h * 2 + (k >> 2)) * (kBM * 128) + (k & 3) * 32);
            const uint64_t db = make_desc_k_sw128(y_addr + (s * 2 + (k >> 2)) * (kBN * 128) + (k & 3) * 32);
            if (leader) umma_bf16(tmem + (h * 2 + buf) * kBN, da, db, idesc_s, k > 0);
          }
        if (leader) umma_commit(&sm.s_full[buf]);
      };
      mbar_wait(&sm.x_full, 0);
      issue_s(0);
      if (n_iter > 1) issue_s(1);
      for (int i = 0; i < n_iter; ++i) {
        const int s = i % kStages, pb = i & 1;
        mbar_wait(&sm.p_full[pb], (i >> 1) & 1);                     // P(i) of both halves is in tensor memory
        tc_fence_after_sync();
#pragma unroll
        for (int h = 0; h < kHalves; ++h)
#pragma unroll
          for (int k = 0; k < 4; ++k) {                              // K steps 0,1: column group 0's P; 2,3: group 1's
            const uint64_t dv = make_desc_k_sw128(v_addr + s * (kBM * 128) + k * 32);
            if (leader)
              umma_bf16_ts(tmem + kTmemO + h * kBM, tmem + (h * 2 + pb) * kBN + (k >> 1) * 32 + (k & 1) * 8, dv, idesc_o,
                           (i > 0) || (k > 0));
          }
        if (leader) umma_commit(&sm.empty[s]);                       // Y_i and V_i consumed
        if (i + 2 < n_iter) issue_s(i + 2);                          // overwrites S/P(i): ordered behind the product above
      }
      if (leader) umma_commit(&sm.o_full);
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;                                    // TMEM lane quarter this warp may access
    const int half = ew >> 3;                                        // stationary half
    const int cg = (ew >> 2) & 1;                                    // 32-column group of the 64-wide S tile
    const int row = quarter * 32 + lane;                             // TMEM lane
    const int mrow = m0 + half * kBM + row;                          // stationary row
    // per-lane constants
    float laneL = 0.f;            // pass A: lse_q * log2e          pass B: b_j * log2e
    int laneI = -1;               // pass A: y_q                    pass B: global item id
    float lane_g = 0.f;           // pass A: g_q
    if (!kPassB) {
      if (mrow < a.Q) {
        laneL = a.lseL[mrow];
        laneI = a.yq[mrow];
        lane_g = a.gq[mrow];
      }
    } else if (mrow < a.n_items) {
      laneL = a.bL[mrow];
      laneI = a.n0 + mrow;
    }
    float db_acc = 0.f;           // pass B: db_j; pass F: l_q (this warp's 32 of every 64 columns)
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    // colA / colB / colI are consecutive [kStages][kBN] arrays: shared-window address of this warp's first column
    const uint32_t col_base = smem_u32(&sm.colA[0][0]) + cg * kColsPerWarp * 4;
    static_assert(offsetof(BwdSmem, colB) == offsetof(BwdSmem, colA) + kStages * kBN * 4 &&
                  offsetof(BwdSmem, colI) == offsetof(BwdSmem, colB) + kStages * kBN * 4, "column vectors must be contiguous");

    for (int i = 0; i < n_iter; ++i) {
      const int buf = i & 1, st = i % kStages;
      const int j0 = (t_begin + i) * kBN;
      const uint32_t sp = lane_base + (half * 2 + buf) * kBN + cg * kColsPerWarp;    // this warp's S columns = its P columns
      mbar_wait(&sm.full[st], (i / kStages) & 1);                    // the column vectors of this tile have landed
      mbar_wait(&sm.s_full[buf], (i >> 1) & 1);
      tc_fence_after_sync();
      uint32_t r[kColsPerWarp];
      tmem_ld_32x32(sp, r);
      tmem_ld_wait(r);
      uint32_t pk[kColsPerWarp / 2];
      const int tgt = kPassB ? 0 : laneI - (a.n0 + j0 + cg * kColsPerWarp);   // pass A: tile-local column of this row's target
#pragma unroll
      for (int u = 0; u < kColsPerWarp; u += 4) {                     // 4 columns per 128-bit read of the column vectors
        // explicit ld.shared: through the generic `sm` reference these compile to LD.E (generic loads, 20% of the stalls)
        const float4 cA = lds_f4(col_base + (st * kBN + u) * 4);
        float4 cB = make_float4(0.f, 0.f, 0.f, 0.f), cI = cB;
        if (kPassB) {
          cB = lds_f4(col_base + ((kStages + st) * kBN + u) * 4);
          cI = lds_f4(col_base + ((2 * kStages + st) * kBN + u) * 4);
        }
        const float av[4] = {cA.x, cA.y, cA.z, cA.w};
        const float bv[4] = {cB.x, cB.y, cB.z, cB.w};
        const int iv[4] = {__float_as_int(cI.x), __float_as_int(cI.y), __float_as_int(cI.z), __float_as_int(cI.w)};
        float pv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float z = __uint_as_float(r[u + e]);
          if (kPassF) {
            pv[e] = ex2_approx(fmaf(z, kLog2e, av[e] - laneL));      // exp(z_j - z_y); the padding columns carry b = -inf
            db_acc += pv[e];
          } else if (!kPassB) {
            pv[e] = ex2_approx(fmaf(z, kLog2e, av[e] - laneL)) - ((u + e == tgt) ? 1.f : 0.f);
          } else {
            pv[e] = bv[e] * (ex2_approx(fmaf(z, kLog2e, laneL - av[e])) - ((iv[e] == laneI) ? 1.f : 0.f));
            db_acc += pv[e];
          }
        }
        pk[u >> 1] = pack_bf16x2(pv[0], pv[1]);
        pk[(u >> 1) + 1] = pack_bf16x2(pv[2], pv[3]);
      }
      tmem_st_32x16(sp, pk);                                         // P over the first half of the columns just read
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.p_full[buf]);
    }

    // ---- O -> global ---------------------------------------------------------------------------------------------
    mbar_wait(&sm.o_full, 0);
    tc_fence_after_sync();
    const bool row_ok = kPassB ? (mrow < a.n_items) : (mrow < a.Q);
    const float scale = (kPassB || kPassF) ? 1.f : lane_g;
#pragma unroll
    for (int c = 0; c < 64; c += 32) {                               // 128 output columns / 2 column groups
      uint32_t r[32];
      tmem_ld_32x32(lane_base + kTmemO + half * kBM + cg * 64 + c, r);
      tmem_ld_wait(r);
      if (row_ok) {
        float* dst = a.out + (long long)mrow * kDim + cg * 64 + c;
#pragma unroll
        for (int u = 0; u < 32; u += 4)
          atomicAdd(reinterpret_cast<float4*>(dst + u),
                    make_float4(__uint_as_float(r[u]) * scale, __uint_as_float(r[u + 1]) * scale,
                                __uint_as_float(r[u + 2]) * scale, __uint_as_float(r[u + 3]) * scale));
      }
    }
    if (kPassB && a.d_b && row_ok) atomicAdd(a.d_b + mrow, db_acc);
    if (kPassF && row_ok) atomicAdd(a.lsum + mrow, db_acc);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem);
  }
}

__global__ void k4_bwd_prep_kernel(const float* __restrict__ b_out, int n_items, int n64, const float* __restrict__ loss_row,
                                   const float* __restrict__ zy, const float* __restrict__ g_row,
                                   const int* __restrict__ y_id, int Q, int q64, float* __restrict__ bL,
                                   float* __restrict__ lseL, float* __restrict__ gq, int* __restrict__ yq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n64) bL[i] = (i < n_items) ? b_out[i] * kLog2e : -INFINITY;
  if (i < q64) {
    const bool ok = i < Q;
    lseL[i] = ok ? (loss_row[i] + zy[i]) * kLog2e : INFINITY;
    gq[i] = ok ? g_row[i] : 0.f;
    yq[i] = ok ? y_id[i] : -2;
  }
}

// pass F bookkeeping: bL as above; zyL [ceil64(Q)] = z_y * log2e (0 padding); yq (-2 padding)
__global__ void k4_fused_prep_kernel(const float* __restrict__ b_out, int n_items, int n64, const float* __restrict__ zy,
                                     const int* __restrict__ y_id, int Q, int q64, float* __restrict__ bL,
                                     float* __restrict__ zyL, int* __restrict__ yq, float* __restrict__ lsum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n64) bL[i] = (i < n_items) ? b_out[i] * kLog2e : -INFINITY;
  if (i < q64) {
    const bool ok = i < Q;
    zyL[i] = ok ? zy[i] * kLog2e : 0.f;
    yq[i] = ok ? y_id[i] : -2;
    lsum[i] = 0.f;
  }
}

// After pass F: d_hout holds O_q = sum_j exp(z_j - z_y) w_j and lsum l_q = sum_j exp(z_j - z_y).  One warp per row:
//     loss_q = log l_q,    dHout[q] = g_q (O_q / l_q - w_{y_q})
// A row whose sum left the fp32 range (some logit more than ~69 nats above the target's) is redone here exactly, with a
// running maximum, on the same bf16 operands (lane per item for the logits, lane per 4 channels for the weighted sum).
constexpr int kFinishWarps = 8;
__global__ void __launch_bounds__(32 * kFinishWarps)
ce_fused_finish_kernel(int Q, const float* __restrict__ lsum, const float* __restrict__ zy, const float* __restrict__ g_row,
                       const int* __restrict__ y_id, const __nv_bfloat16* __restrict__ hout,
                       const __nv_bfloat16* __restrict__ w_out_t, const float* __restrict__ b_out, int n_items, int n0,
                       float* __restrict__ d_hout, float* __restrict__ loss_row) {
  __shared__ float hs[kFinishWarps][kDim];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long q = (long long)blockIdx.x * kFinishWarps + wib;
  if (q >= Q) return;
  const float l = lsum[q];
  float4 o = reinterpret_cast<const float4*>(d_hout + q * kDim)[lane];
  float loss;
  if (l < 1.2676506e30f && l > 0.f) {                         // 2^100: the sums are in range
    loss = __logf(l);
    const float inv = 1.f / l;
    o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
  } else {
    const uint2 hq = reinterpret_cast<const uint2*>(hout + q * kDim)[lane];
    hs[wib][4 * lane + 0] = bf16_lo(hq.x); hs[wib][4 * lane + 1] = bf16_hi(hq.x);
    hs[wib][4 * lane + 2] = bf16_lo(hq.y); hs[wib][4 * lane + 3] = bf16_hi(hq.y);
    __syncwarp();
    auto logit = [&](int j) {
      const uint4* wr = reinterpret_cast<const uint4*>(w_out_t + (long long)j * kWtPitchBf16);
      float z = b_out[j];
#pragma unroll 4
      for (int c = 0; c < 16; ++c) {
        const uint4 w8 = __ldg(wr + c);
        const float* h = &hs[wib][c * 8];
        z = fmaf(h[0], bf16_lo(w8.x), z); z = fmaf(h[1], bf16_hi(w8.x), z); z = fmaf(h[2], bf16_lo(w8.y), z);
        z = fmaf(h[3], bf16_hi(w8.y), z); z = fmaf(h[4], bf16_lo(w8.z), z); z = fmaf(h[5], bf16_hi(w8.z), z);
        z = fmaf(h[6], bf16_lo(w8.w), z); z = fmaf(h[7], bf16_hi(w8.w), z);
      }
      return z;
    };
    float m = -INFINITY;
    for (int j = lane; j < n_items; j += 32) m = fmaxf(m, logit(j));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    float ls = 0.f;
    o = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = 0; j0 < n_items; j0 += 32) {
      const int j = j0 + lane;
      const float p = j < n_items ? expf(logit(j) - m) : 0.f;
      ls += p;
      const int cnt = n_items - j0 < 32 ? n_items - j0 : 32;
      for (int i = 0; i < cnt; ++i) {
        const float pj = __shfl_sync(0xffffffffu, p, i);
        const uint2 w4 = __ldg(reinterpret_cast<const uint2*>(w_out_t + (long long)(j0 + i) * kWtPitchBf16) + lane);
        o.x = fmaf(pj, bf16_lo(w4.x), o.x); o.y = fmaf(pj, bf16_hi(w4.x), o.y);
        o.z = fmaf(pj, bf16_lo(w4.y), o.z); o.w = fmaf(pj, bf16_hi(w4.y), o.w);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, d);
    loss = m + logf(ls) - zy[q];
    const float inv = 1.f / ls;
    o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
  }
  const int y = y_id[q] - n0;
  float4 wy = make_float4(0.f, 0.f, 0.f, 0.f);
  if (y >= 0 && y < n_items) {
    const uint2 w4 = __ldg(reinterpret_cast<const uint2*>(w_out_t + (long long)y * kWtPitchBf16) + lane);
    wy = make_float4(bf16_lo(w4.x), bf16_hi(w4.x), bf16_lo(w4.y), bf16_hi(w4.y));
  }
  const float g = g_row[q];
  reinterpret_cast<float4*>(d_hout + q * kDim)[lane] = make_float4(g * (o.x - wy.x), g * (o.y - wy.y), g * (o.z - wy.z), g * (o.w - wy.w));
  if (lane == 0) loss_row[q] = loss;
}

// src [R,128] f32 -> rows [R,128] bf16 (optional) and its transpose [128, r_pad] bf16 (optional), 32x32 smem tiles
__global__ void cast_transpose_bf16_kernel(const void* __restrict__ src, int src_bf16, long long R,
                                           __nv_bfloat16* __restrict__ rows, __nv_bfloat16* __restrict__ tr, long long r_pad) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;       // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const long long r = r0 + i;
    const float v = (r >= R) ? 0.f
                    : src_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[r * kDim + c0 + tx])
                               : reinterpret_cast<const float*>(src)[r * kDim + c0 + tx];
    tile[i][tx] = v;
    if (rows && r < R) rows[r * kDim + c0 + tx] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  if (tr) {
    for (int i = ty; i < 32; i += 8) {
      const long long r = r0 + tx;
      if (r < r_pad) tr[(long long)(c0 + i) * r_pad + r] = __float2bfloat16_rn(tile[tx][i]);
    }
  }
}

template <int kMode>
int32_t launch_bwd(const void* x, uint64_t x_rows, uint32_t x_pitch, const void* y, uint64_t y_rows, uint32_t y_pitch,
                   const void* v, uint64_t v_cols, uint64_t v_pitch, BwdArgs a, cudaStream_t st) {
  CUtensorMap tx, ty, tv;
  int32_t rc = make_tmap_bf16(&tx, x, x_rows, kDim, x_pitch, 64, kBM, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&ty, y, y_rows, kDim, y_pitch, 64, kBN, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tv, v, kDim, (uint32_t)v_cols, (uint32_t)v_pitch, 64, kBM, 128);
  if (rc) return rc;
  const int m_tiles = ceil_div((long long)x_rows, kHalves * kBM);
  const int s_tiles = ceil_div(a.n_stream, kBN);
  // about four waves of CTAs (1 CTA per SM); among the split counts around that, the one whose last wave is fullest
  // (config 5's pass B: 82 item tiles x 8 splits = 4.43 waves ran as 5; 9 splits = 4.99)
  int ns = 1;
  if (m_tiles < 4 * 148) {
    const int lo = ceil_div(3 * 148, m_tiles), hi = ceil_div(6 * 148, m_tiles);
    double best = -1.0;
    for (int c = lo; c <= hi; ++c) {
      const long long ctas = (long long)m_tiles * c;
      const double eff = (double)ctas / (double)(ceil_div(ctas, 148) * 148);
      if (eff > best + 1e-9) { best = eff; ns = c; }
    }
  }
  if (ns > s_tiles) ns = s_tiles;
  if (ns > 65535) ns = 65535;
  if (ns < 1) ns = 1;
  a.n_split = ns;
  const size_t smem = sizeof(BwdSmem) + 1024;
  auto kern = k4_ce_backward_bf16<kMode>;
  HTCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(m_tiles, ns), kThreads, smem, st>>>(tx, ty, tv, a);
  HTCN_LAUNCH_CHECK("k4_ce_backward_bf16");
  return HTCN_OK;
}

}  // namespace
}  // namespace htcn

using namespace htcn;

extern "C" int32_t htcn_cast_transpose_bf16(const void* src, int32_t src_dtype, int64_t R, void* dst_rows, void* dst_t,
                                            int64_t r_pad, void* stream) {
  HTCN_REQUIRE(src && R > 0 && (dst_rows || dst_t), "cast_transpose_bf16: bad args");
  HTCN_REQUIRE(src_dtype == HTCN_F32 || src_dtype == HTCN_BF16, "cast_transpose_bf16: src_dtype %d", src_dtype);
  HTCN_REQUIRE(!dst_t || (r_pad >= R && r_pad % 8 == 0), "cast_transpose_bf16: r_pad=%lld must be >= R and a multiple of 8",
               (long long)r_pad);
  const long long span = dst_t ? r_pad : R;
  cast_transpose_bf16_kernel<<<dim3(ceil_div(span, 32), kDim / 32), dim3(32, 8), 0, as_stream(stream)>>>(
      src, src_dtype == HTCN_BF16, R, reinterpret_cast<__nv_bfloat16*>(dst_rows), reinterpret_cast<__nv_bfloat16*>(dst_t), r_pad);
  HTCN_LAUNCH_CHECK("cast_transpose_bf16_kernel");
  return HTCN_OK;
}

extern "C" int32_t htcn_score_ce_backward_bf16(const void* hout, const void* hout_t, int64_t q_pad, int32_t Q,
                                               const void* w_out_t, const void* w_out, int64_t n_pad,
                                               const float* b_out, int32_t n_items, int32_t n0, const int32_t* y_id,
                                               const float* loss_row, const float* target_logit, const float* g_row,
                                               float* workspace, float* d_hout, float* d_w_out_t, float* d_b_out,
                                               void* stream) {
  HTCN_REQUIRE(hout && hout_t && w_out_t && w_out && b_out && y_id && loss_row && target_logit && g_row && workspace &&
                   d_hout && d_w_out_t,
               "score_ce_backward_bf16: NULL pointer");
  HTCN_REQUIRE(Q >= 0 && n_items > 0 && q_pad >= Q && q_pad % 8 == 0 && n_pad >= n_items && n_pad % 8 == 0,
               "score_ce_backward_bf16: Q=%d q_pad=%lld n_items=%d n_pad=%lld", Q, (long long)q_pad, n_items, (long long)n_pad);
  if (Q == 0) return HTCN_OK;
  cudaStream_t st = as_stream(stream);
  HTCN_CUDA(cudaMemsetAsync(d_hout, 0, sizeof(float) * (size_t)Q * kDim, st));
  const int q64 = ceil_div(Q, kBN) * kBN, n64 = ceil_div(n_items, kBN) * kBN;
  float* bL = workspace;                 // HTCN_CE_BWD_BF16_WS_FLOATS(Q, n_items)
  float* lseL = bL + n64;
  float* gq = lseL + q64;
  int* yq = reinterpret_cast<int*>(gq + q64);
  k4_bwd_prep_kernel<<<ceil_div(q64 > n64 ? q64 : n64, 256), 256, 0, st>>>(b_out, n_items, n64, loss_row, target_logit, g_row,
                                                                          y_id, Q, q64, bL, lseL, gq, yq);
  HTCN_LAUNCH_CHECK("k4_bwd_prep_kernel");
  BwdArgs a{bL, lseL, gq, yq, d_hout, nullptr, nullptr, Q, n_items, n0, n_items, 1};
  int32_t rc = launch_bwd<0>(hout, (uint64_t)Q, kDim, w_out_t, (uint64_t)n_items, kWtPitchBf16, w_out, (uint64_t)n_items,
                                 (uint64_t)n_pad, a, st);
  if (rc) return rc;
  a.out = d_w_out_t;
  a.d_b = d_b_out;
  a.n_stream = Q;
  return launch_bwd<1>(w_out_t, (uint64_t)n_items, kWtPitchBf16, hout, (uint64_t)Q, kDim, hout_t, (uint64_t)Q,
                          (uint64_t)q_pad, a, st);
}


// Forward + backward of the full-catalog softmax-CE head in TWO catalog sweeps (pass F: loss and dHout; pass B: dW^T, db)
extern "C" int32_t htcn_score_ce_fwd_bwd_bf16(const void* hout, const void* hout_t, int64_t q_pad, int32_t Q,
                                              const void* w_out_t, const void* w_out, int64_t n_pad, const float* b_out,
                                              int32_t n_items, int32_t n0, const int32_t* y_id, const float* target_logit,
                                              const float* g_row, float* workspace, float* loss_row, float* d_hout,
                                              float* d_w_out_t, float* d_b_out, void* stream) {
  HTCN_REQUIRE(hout && hout_t && w_out_t && w_out && b_out && y_id && loss_row && target_logit && g_row && workspace &&
                   d_hout && d_w_out_t,
               "score_ce_fwd_bwd_bf16: NULL pointer");
  HTCN_REQUIRE(Q >= 0 && n_items > 0 && q_pad >= Q && q_pad % 8 == 0 && n_pad >= n_items && n_pad % 8 == 0,
               "score_ce_fwd_bwd_bf16: Q=%d q_pad=%lld n_items=%d n_pad=%lld", Q, (long long)q_pad, n_items, (long long)n_pad);
  if (Q == 0) return HTCN_OK;
  cudaStream_t st = as_stream(stream);
  HTCN_CUDA(cudaMemsetAsync(d_hout, 0, sizeof(float) * (size_t)Q * kDim, st));
  const int q64 = ceil_div(Q, kBN) * kBN, n64 = ceil_div(n_items, kBN) * kBN;
  float* bL = workspace;                 // HTCN_CE_BWD_BF16_WS_FLOATS(Q, n_items)
  float* lseL = bL + n64;
  float* gq = lseL + q64;
  int* yq = reinterpret_cast<int*>(gq + q64);
  float* lsum = reinterpret_cast<float*>(yq + q64);
  const int pb = ceil_div(q64 > n64 ? q64 : n64, 256);
  k4_fused_prep_kernel<<<pb, 256, 0, st>>>(b_out, n_items, n64, target_logit, y_id, Q, q64, bL, lseL, yq, lsum);
  HTCN_LAUNCH_CHECK("k4_fused_prep_kernel");
  BwdArgs a{bL, lseL, gq, yq, d_hout, nullptr, lsum, Q, n_items, n0, n_items, 1};
  int32_t rc = launch_bwd<2>(hout, (uint64_t)Q, kDim, w_out_t, (uint64_t)n_items, kWtPitchBf16, w_out, (uint64_t)n_items,
                             (uint64_t)n_pad, a, st);
  if (rc) return rc;
  ce_fused_finish_kernel<<<ceil_div(Q, kFinishWarps), 32 * kFinishWarps, 0, st>>>(
      Q, lsum, target_logit, g_row, y_id, reinterpret_cast<const __nv_bfloat16*>(hout),
      reinterpret_cast<const __nv_bfloat16*>(w_out_t), b_out, n_items, n0, d_hout, loss_row);
  HTCN_LAUNCH_CHECK("ce_fused_finish_kernel");
  k4_bwd_prep_kernel<<<pb, 256, 0, st>>>(b_out, n_items, n64, loss_row, target_logit, g_row, y_id, Q, q64, bL, lseL, gq, yq);
  HTCN_LAUNCH_CHECK("k4_bwd_prep_kernel");
  a.out = d_w_out_t;
  a.d_b = d_b_out;
  a.n_stream = Q;
  return launch_bwd<1>(w_out_t, (uint64_t)n_items, kWtPitchBf16, hout, (uint64_t)Q, kDim, hout_t, (uint64_t)Q,
                       (uint64_t)q_pad, a, st);
}
