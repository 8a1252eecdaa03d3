// K4 (fp32 tier): full-catalog scoring as an FFMA tile GEMM with the CE / rank / top-k epilogues
// fused, so the [Q, N] logits never reach HBM.  Replaces model_tcn.py:41 + loss.py:20-21,179,120.
//
// One CTA keeps a 128-row query tile resident (k-major in shared memory) and sweeps its catalog
// split in tiles of 64 items.  Every logit is accumulated as a single fmaf chain over k = 0..127
// (then + bias), the same chain htcn_target_logit uses, so z[q, y] compares EQUAL to the
// target logit and the strict-greater rank of loss.py:179 is self-consistent.
// The epilogue is run by one thread per row ("lane = row"), the organisation the tcgen05 tier
// gets for free from TMEM lanes.
#include "common.cuh"

namespace htcn {

constexpr int kQM = 128;     // query rows per CTA
constexpr int kQN = 64;      // items per tile
constexpr int kK4Threads = 256;

__global__ void __launch_bounds__(kK4Threads, 1)
k4_score_f32(ScoreArgs a) {
  extern __shared__ __align__(16) float sm[];
  const float* hout = reinterpret_cast<const float*>(a.hout);
  const float* wt = reinterpret_cast<const float*>(a.wt);
  const int P = a.planes > 1 ? a.planes : 1;   // 128-wide planes of the contraction (2: HTCN_F32_W256)
  float* As = sm;                          // [P][128 k][128 rows]
  float* Bs = As + (size_t)P * kDim * kQM; // [128 k][64 items]; reused as Zs[128 rows][64] (xor-swizzled)
  float* bias_s = Bs + kDim * kQN;         // [64]
  float* heap_v = bias_s + kQN;            // [k][128]
  int* heap_i = reinterpret_cast<int*>(heap_v + a.k * kQM);
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * kQM;
  const int split = blockIdx.y;
  // catalog split: contiguous ranges of whole 64-item tiles
  const int n_tiles = (a.n_items + kQN - 1) / kQN;
  const int t_begin = (int)((long long)n_tiles * split / a.n_split);
  const int t_end = (int)((long long)n_tiles * (split + 1) / a.n_split);

  for (int p = 0; p < P; ++p) {  // A tile(s), transposed: As[p][k][row]; plane p of hout starts Q*128 floats further
    const int row = tid & 127, kq = tid >> 7;      // 2 groups of 64 k
    const bool ok = q0 + row < a.Q;
    const float4* src = reinterpret_cast<const float4*>(hout + ((long long)p * a.Q + q0 + row) * kDim + kq * 64);
    float* Ap = As + (size_t)p * kDim * kQM;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const float4 v = ok ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      const int k = kq * 64 + i * 4;
      Ap[(k + 0) * kQM + row] = v.x; Ap[(k + 1) * kQM + row] = v.y;
      Ap[(k + 2) * kQM + row] = v.z; Ap[(k + 3) * kQM + row] = v.w;
    }
  }
  // per-row epilogue state (threads 0..127 own row tid)
  const bool row_thread = tid < kQM;
  const bool row_ok = row_thread && (q0 + tid < a.Q);
  float zy = 0.f, run_m = -INFINITY, run_s = 0.f, thr = -INFINITY;
  const float rs = (row_ok && a.row_scale) ? a.row_scale[q0 + tid] : 1.f;
  int cnt = 0;
  RowHeap heap{heap_v + tid, heap_i + tid, kQM, a.k};
  if (row_ok && (a.flags & (HTCN_SCORE_CE | HTCN_SCORE_RANK))) zy = a.zy[q0 + tid];
  if (row_thread && (a.flags & HTCN_SCORE_TOPK)) heap.init();

  const int tx = tid & 15, ty = tid >> 4;    // micro-tile: rows {ty*4.., 64+ty*4..} x cols tx*4..+3
  for (int t = t_begin; t < t_end; ++t) {
    const int j0 = t * kQN;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int p = 0; p < P; ++p) {             // one fmaf chain per logit over k = 0 .. 128 P - 1
      __syncthreads();                        // previous tile's epilogue / previous plane's math is done with Bs/Zs
      {  // B tile, transposed: Bs[k][item]
        const int item = tid & 63, kq = tid >> 6;    // 4 groups of 32 k
        const bool ok = j0 + item < a.n_items;
        const float4* src = reinterpret_cast<const float4*>(wt + ((long long)(j0 + item) * P + p) * kDim + kq * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 v = ok ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          const int k = kq * 32 + i * 4;
          Bs[(k + 0) * kQN + item] = v.x; Bs[(k + 1) * kQN + item] = v.y;
          Bs[(k + 2) * kQN + item] = v.z; Bs[(k + 3) * kQN + item] = v.w;
        }
        if (p == 0 && tid < kQN) bias_s[tid] = (j0 + tid < a.n_items) ? __ldg(a.b_out + j0 + tid) : 0.f;
      }
      __syncthreads();
      const float* Ap = As + (size_t)p * kDim * kQM;
#pragma unroll 8
      for (int k = 0; k < kDim; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(Ap + k * kQM + ty * 4);
        const float4 a1 = *reinterpret_cast<const float4*>(Ap + k * kQM + 64 + ty * 4);
        const float4 b = *reinterpret_cast<const float4*>(Bs + k * kQN + tx * 4);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    __syncthreads();                          // everyone is done reading Bs -> becomes Zs
    float* Zs = Bs;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) Zs[r * kQN + ((tx * 4 + j) ^ (r & 31))] = acc[i][j];
    }
    __syncthreads();
    if (row_ok) {
      const int lim = min(kQN, a.n_items - j0);
      const int r = tid;
      for (int c = 0; c < lim; ++c) {
        const float z = Zs[r * kQN + (c ^ (r & 31))] + bias_s[c];
        if (a.flags & HTCN_SCORE_CE) {        // online softmax partial (max, sum exp(z - max)) of the (scaled) logits
          const float zs = a.row_scale ? z * rs : z;
          if (zs > run_m) {
            run_s = run_s * expf(run_m - zs) + 1.0f;
            run_m = zs;
          } else {
            run_s += expf(zs - run_m);
          }
        }
        if (a.flags & HTCN_SCORE_RANK) cnt += (z > zy) ? 1 : 0;
        if ((a.flags & HTCN_SCORE_TOPK) && z > thr) thr = heap.replace_root(z, a.n0 + j0 + c);
      }
    }
  }
  if (row_ok) {
    const long long o = (long long)split * a.Q + q0 + tid;
    if (a.flags & HTCN_SCORE_CE) {
      a.part_max[o] = run_m;
      a.part_sum[o] = run_s;
    }
    if (a.flags & HTCN_SCORE_RANK) a.part_cnt[o] = cnt;
    if (a.flags & HTCN_SCORE_TOPK) {
      for (int s = 0; s < a.k; ++s) {
        const int id = heap.i(s);
        a.topk_val[o * a.k + s] = heap.v(s);
        a.topk_idx[o * a.k + s] = (id == 0x7fffffff) ? -1 : id;
      }
    }
  }
}

// zy[q] = fmaf-chain_k(hout[q,k], wt[y-n0,k]) + b[y-n0]  -- identical arithmetic to the sweep above
__global__ void k4_target_logit_f32(const float* __restrict__ hout, const float* __restrict__ wt,
                                    const float* __restrict__ b_out, const int* __restrict__ y_id, int Q,
                                    int n_items, int n0, float* __restrict__ zy, int P) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const int y = y_id[q] - n0;
  if (y < 0 || y >= n_items) return;
  float acc = 0.f;
  for (int p = 0; p < P; ++p) {               // planes of the block-planar hout; wt rows are P*128 floats
    const float4* h = reinterpret_cast<const float4*>(hout + ((long long)p * Q + q) * kDim);
    const float4* w = reinterpret_cast<const float4*>(wt + ((long long)y * P + p) * kDim);
#pragma unroll 8
    for (int i = 0; i < kDim / 4; ++i) {
      const float4 a = __ldg(h + i), b = __ldg(w + i);
      acc = fmaf(a.x, b.x, acc);
      acc = fmaf(a.y, b.y, acc);
      acc = fmaf(a.z, b.z, acc);
      acc = fmaf(a.w, b.w, acc);
    }
  }
  zy[q] = acc + __ldg(b_out + y);
}

int32_t score_f32(const ScoreArgs& a, cudaStream_t st) {
  const int P = a.planes > 1 ? a.planes : 1;
  const size_t smem = sizeof(float) * ((size_t)P * kDim * kQM + kDim * kQN + kQN) +
                      ((a.flags & HTCN_SCORE_TOPK) ? (size_t)a.k * kQM * 8 : 0);
  if (smem > 227 * 1024) {
    set_error("score(f32): %d planes with k=%d need %zu B of shared memory (max 232448); use k <= 64 with 256-wide embeddings",
              P, a.k, smem);
    return HTCN_ERR_UNSUPPORTED;
  }
  HTCN_CUDA(cudaFuncSetAttribute(k4_score_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(a.Q, kQM), a.n_split);
  k4_score_f32<<<grid, kK4Threads, smem, st>>>(a);
  HTCN_LAUNCH_CHECK("k4_score_f32");
  return HTCN_OK;
}

int32_t target_logit_f32(const float* hout, const float* wt, const float* b_out, const int* y_id, int Q,
                         int n_items, int n0, float* zy, cudaStream_t st, int planes) {
  k4_target_logit_f32<<<ceil_div(Q, 128), 128, 0, st>>>(hout, wt, b_out, y_id, Q, n_items, n0, zy, planes > 1 ? planes : 1);
  HTCN_LAUNCH_CHECK("k4_target_logit_f32");
  return HTCN_OK;
}

}  // namespace htcn
