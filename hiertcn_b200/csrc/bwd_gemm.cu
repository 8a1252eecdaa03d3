// fp32 GEMM building blocks of the backward pass (see train.cuh).  128x128 output tiles, 256 threads, 8x8 register
// tiles, operands staged through shared memory in 16-deep slices -- the same inner loop as the fp32 conv level
// (k2_tcn_f32.cu).  These carry < 2% of a training step's FLOPs (the catalog products are in bwd_k4.cu).
#include "train.cuh"

namespace htcn {

namespace {
constexpr int kT = 128;       // tile edge
constexpr int kS = 16;        // contraction slice
constexpr int kThreads = 256;

__device__ __forceinline__ void fma_tile(float (&acc)[8][8], const float* __restrict__ a_row, const float* __restrict__ b_row,
                                         int ay, int bx) {
  const float4 a0 = *reinterpret_cast<const float4*>(a_row + ay * 4);
  const float4 a1 = *reinterpret_cast<const float4*>(a_row + 64 + ay * 4);
  const float4 b0 = *reinterpret_cast<const float4*>(b_row + bx * 4);
  const float4 b1 = *reinterpret_cast<const float4*>(b_row + 64 + bx * 4);
  const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
}
}  // namespace

// ---- C[M,N] (+)= A[M,K] * op(B) -------------------------------------------------------------------------------
template <bool kTransB>
__global__ void __launch_bounds__(kThreads)
sgemm_kernel(long long M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
             float* __restrict__ C, int ldc, int accumulate) {
  __shared__ __align__(16) float As[kS][kT];     // As[kk][row]
  __shared__ __align__(16) float Bs[kS][kT];     // Bs[kk][col]
  const int tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * kT;
  const int n0 = blockIdx.y * kT;
  const int tx = tid & 15, ty = tid >> 4;
  const int lrow = tid & 127, lhalf = tid >> 7;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += kS) {
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    if (r0 + lrow < M) {
      const float4* p = reinterpret_cast<const float4*>(A + (r0 + lrow) * lda + k0 + lhalf * 8);
      a0 = p[0];
      a1 = p[1];
    }
    float4 b0, b1;
    if (kTransB) {          // B[N,K]: row n0 + lrow, 8 consecutive k
      const float4* p = reinterpret_cast<const float4*>(B + (long long)(n0 + lrow) * ldb + k0 + lhalf * 8);
      b0 = __ldg(p);
      b1 = __ldg(p + 1);
    } else {                // B[K,N]: 16 rows x 128 cols = 512 float4, two per thread
      const int e0 = tid, e1 = tid + 256;
      b0 = __ldg(reinterpret_cast<const float4*>(B + (long long)(k0 + (e0 >> 5)) * ldb + n0 + (e0 & 31) * 4));
      b1 = __ldg(reinterpret_cast<const float4*>(B + (long long)(k0 + (e1 >> 5)) * ldb + n0 + (e1 & 31) * 4));
    }
    __syncthreads();
    {
      const float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) As[lhalf * 8 + i][lrow] = v[i];
    }
    if (kTransB) {
      const float v[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) Bs[lhalf * 8 + i][lrow] = v[i];
    } else {
      reinterpret_cast<float4*>(&Bs[0][0])[tid] = b0;
      reinterpret_cast<float4*>(&Bs[0][0])[tid + 256] = b1;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kS; ++kk) fma_tile(acc, &As[kk][0], &Bs[kk][0], ty, tx);
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int lr = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4);
    const long long r = r0 + lr;
    if (r >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      float4* dst = reinterpret_cast<float4*>(C + r * ldc + n0 + jh * 64 + tx * 4);
      float4 o = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      if (accumulate) {
        const float4 c = *dst;
        o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
      }
      *dst = o;
    }
  }
}

int32_t sgemm(bool trans_b, long long M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
              int ldc, bool accumulate, cudaStream_t st) {
  HTCN_REQUIRE(M > 0 && N > 0 && N % kT == 0 && K > 0 && K % kS == 0, "sgemm: M=%lld N=%d K=%d", M, N, K);
  HTCN_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0, "sgemm: leading dimensions must be multiples of 4");
  dim3 grid(ceil_div(M, kT), N / kT);
  if (trans_b)
    sgemm_kernel<true><<<grid, kThreads, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate ? 1 : 0);
  else
    sgemm_kernel<false><<<grid, kThreads, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate ? 1 : 0);
  HTCN_LAUNCH_CHECK("sgemm_kernel");
  return HTCN_OK;
}

// ---- C[c,f] += sum_r A[src(r), c] * D[r, f]  (split over row chunks, atomic accumulation) -----------------------
__global__ void __launch_bounds__(kThreads)
sgemm_tn_atomic_kernel(long long R, int chunk, const void* __restrict__ A, int lda, int a_bf16,
                       const float* __restrict__ D, int ldd, float* __restrict__ C, int ldc, int shift, int T,
                       SlotTable slots) {
  __shared__ __align__(16) float As[kS][kT];     // As[rr][c]
  __shared__ __align__(16) float Ds[kS][kT];     // Ds[rr][f]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long c_begin = (long long)blockIdx.x * chunk;
  const long long c_end = (c_begin + chunk < R) ? c_begin + chunk : R;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (long long rb = c_begin; rb < c_end; rb += kS) {
    float4 a[2], d[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int e = tid + h * 256;
      const long long r = rb + (e >> 5);
      const int col = (e & 31) * 4;
      a[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      d[h] = a[h];
      if (r < c_end) {
        d[h] = *reinterpret_cast<const float4*>(D + r * ldd + col);
        bool ok = true;
        if (shift > 0) {
          const int p = (int)(r % T);
          int s = 0;
          while (s + 1 < slots.n && slots.off[s + 1] <= p) ++s;
          ok = (p - slots.off[s] - shift >= 0);
        }
        if (ok) {
          if (a_bf16) {
            const uint2 q = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(A) + (r - shift) * lda + col);
            a[h] = make_float4(bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y));
          } else {
            a[h] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(A) + (r - shift) * lda + col);
          }
        }
      }
    }
    __syncthreads();
    reinterpret_cast<float4*>(&As[0][0])[tid] = a[0];
    reinterpret_cast<float4*>(&As[0][0])[tid + 256] = a[1];
    reinterpret_cast<float4*>(&Ds[0][0])[tid] = d[0];
    reinterpret_cast<float4*>(&Ds[0][0])[tid + 256] = d[1];
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kS; ++rr) fma_tile(acc, &As[rr][0], &Ds[rr][0], ty, tx);
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4);
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      float* dst = C + (long long)c * ldc + jh * 64 + tx * 4;
      atomicAdd(reinterpret_cast<float4*>(dst),
                make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]));
    }
  }
}

int32_t sgemm_tn_atomic(long long R, const void* A, int lda, const float* D, int ldd, float* C, int ldc, int shift,
                        int T, const SlotTable* slots, cudaStream_t st, bool a_bf16) {
  HTCN_REQUIRE(R > 0 && lda % 4 == 0 && ldd % 4 == 0 && ldc % 4 == 0, "sgemm_tn_atomic: R=%lld", R);
  HTCN_REQUIRE(shift == 0 || (slots && T > 0), "sgemm_tn_atomic: a shifted product needs the slot table");
  long long chunk = (R + 295) / 296;                  // about two waves of CTAs
  chunk = ((chunk + kS - 1) / kS) * kS;
  if (chunk < 128) chunk = 128;
  if (chunk > 8192) chunk = 8192;
  SlotTable none{};
  sgemm_tn_atomic_kernel<<<ceil_div(R, chunk), kThreads, 0, st>>>(R, (int)chunk, A, lda, a_bf16 ? 1 : 0, D, ldd, C, ldc,
                                                                  shift, T, slots ? *slots : none);
  HTCN_LAUNCH_CHECK("sgemm_tn_atomic_kernel");
  return HTCN_OK;
}

// ---- out[f] += sum_r D[r,f] ------------------------------------------------------------------------------------
__global__ void colsum_atomic_kernel(long long R, int chunk, const float* __restrict__ D, int ldd, int cols,
                                     float* __restrict__ out) {
  const int f = threadIdx.x;
  if (f >= cols) return;
  const long long b = (long long)blockIdx.x * chunk;
  const long long e = (b + chunk < R) ? b + chunk : R;
  float s = 0.f;
  for (long long r = b; r < e; ++r) s += D[r * ldd + f];
  atomicAdd(out + f, s);
}

int32_t colsum_atomic(long long R, const float* D, int ldd, int cols, float* out, cudaStream_t st) {
  HTCN_REQUIRE(R > 0 && cols > 0 && cols <= 256, "colsum_atomic: R=%lld cols=%d", R, cols);
  long long chunk = (R + 591) / 592;
  if (chunk < 64) chunk = 64;
  colsum_atomic_kernel<<<ceil_div(R, chunk), 256, 0, st>>>(R, (int)chunk, D, ldd, cols, out);
  HTCN_LAUNCH_CHECK("colsum_atomic_kernel");
  return HTCN_OK;
}

}  // namespace htcn
