// Shared host/device helpers for libhtcn (sm_100a).  See include/htcn.h for the boundary.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

#include "../../include/htcn.h"

namespace htcn {

constexpr int kDim = HTCN_DIM;  // D = C = H = 128
// bf16 tier: W_out^T rows are stored augmented, [128 weights | b_hi | b_lo | 14 zeros], so that the tensor core
// adds the bias (htcn_prepare_wout writes this layout; the fp32 tier keeps plain 128-float rows + b_out)
constexpr int kWtPitchBf16 = HTCN_WT_PITCH_BF16;

// thread-local error message (htcn_last_error)
void set_error(const char* fmt, ...);
int32_t cuda_fail(cudaError_t e, const char* what);

struct SlotTable {  // passed by value to kernels: re-entrant, no __constant__ state
  int32_t n;                       // S
  int32_t off[HTCN_MAX_SLOTS + 1];
};

// arguments of one catalog-scoring sweep (K4); hout / wt are f32 or bf16 depending on the tier
struct ScoreArgs {
  const void* hout;          // [Q,128]
  const void* wt;            // [n_items,128]  W_out^T shard
  const float* b_out;        // [n_items]
  const int* y_id;           // [Q] global ids or NULL
  const float* zy;           // [Q] target logits
  float* part_max; float* part_sum; int* part_cnt;   // [n_split,Q]
  float* topk_val; int* topk_idx;                    // [n_split,Q,k]
  int Q, n_items, n0, k, n_split;
  unsigned flags;
  int planes;                // fp32 tier only, 0/1 = plain; 2: hout is [2][Q][128] (block-planar), wt rows are 256 floats
                             // (HTCN_F32_W256: a 256-channel last level, args.py:310-311)
  const float* row_scale;    // [Q] or NULL: the CE sum runs on row_scale[q] * z (l2-normalised head, model_tcn.py:42-43);
                             // part_max then holds the SCALED reference point; ranks always compare the raw logits
  const int* fold_self;      // bf16 folded CE + rank sweep only (k4_score_bf16.cu, kFold): [Q] sign bit of the target column's
                             // own accumulator (taken out of the count by split 0); hout then holds the NEGATED embeddings
};

#define HTCN_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::htcn::set_error(__VA_ARGS__);      \
      return HTCN_ERR_INVALID;             \
    }                                      \
  } while (0)

#define HTCN_CUDA(call)                                        \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) return ::htcn::cuda_fail(e__, #call); \
  } while (0)

#define HTCN_LAUNCH_CHECK(name)                                   \
  do {                                                            \
    cudaError_t e__ = cudaGetLastError();                         \
    if (e__ != cudaSuccess) return ::htcn::cuda_fail(e__, name);  \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_na_f4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_na_u2(uint2* p, const uint2& v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits)
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// ---- per-row top-k min-heap living in shared memory, one heap per thread ("lane = row") ------
// Layout: val[slot * stride + row], idx[...]: all lanes touching the same slot hit distinct banks.
// Order: an entry is WORSE if its value is smaller, or equal with a larger index; the root is the
// worst kept entry.  Scans feed ascending indices, so a candidate must be strictly greater than the
// root value to displace it (tf.nn.top_k keeps the lower index among equals).
struct RowHeap {
  float* val;
  int* idx;
  int stride;
  int k;
  __device__ __forceinline__ float& v(int s) { return val[s * stride]; }
  __device__ __forceinline__ int& i(int s) { return idx[s * stride]; }
  __device__ __forceinline__ void init() {
    for (int s = 0; s < k; ++s) {
      v(s) = -INFINITY;
      i(s) = 0x7fffffff;
    }
  }
  static __device__ __forceinline__ bool worse(float va, int ia, float vb, int ib) {
    return va < vb || (va == vb && ia > ib);
  }
  // replace the root with (z, j) and restore the heap; returns the new root value (threshold)
  __device__ __forceinline__ float replace_root(float z, int j) {
    int p = 0;
    while (true) {
      int l = 2 * p + 1;
      if (l >= k) break;
      int r = l + 1;
      float vl = v(l);
      int il = i(l);
      int c = l;
      float vc = vl;
      int ic = il;
      if (r < k) {
        float vr = v(r);
        int ir = i(r);
        if (worse(vr, ir, vl, il)) {
          c = r;
          vc = vr;
          ic = ir;
        }
      }
      if (worse(vc, ic, z, j)) {
        v(p) = vc;
        i(p) = ic;
        p = c;
      } else {
        break;
      }
    }
    v(p) = z;
    i(p) = j;
    return v(0);
  }
};

// ---- (score desc, index asc) = tf.nn.top_k order (loss.py:120) as ONE descending 64-bit key ---------------------------
// high word: order-preserving image of the score; low word: ~index.  Key 0 = empty slot (sorts last).
__device__ __forceinline__ unsigned long long topk_key(float v, int idx) {
  if (v == 0.f) v = 0.f;                                     // -0.0 and +0.0 compare equal: one key for both
  const uint32_t u = __float_as_uint(v);
  const uint32_t ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)ord << 32) | (uint32_t)(~idx);
}
__device__ __forceinline__ float topk_key_val(unsigned long long key) {
  const uint32_t ord = (uint32_t)(key >> 32);
  return __uint_as_float((ord & 0x80000000u) ? (ord & 0x7FFFFFFFu) : ~ord);
}
__device__ __forceinline__ int topk_key_idx(unsigned long long key) { return (int)(~(uint32_t)key); }

// block-wide bitonic sort of n_pad (power of two) keys in shared memory, DESCENDING; every thread of the CTA calls it
// (ends with a __syncthreads)
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, int n_pad) {
  for (int size = 2; size <= n_pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (n_pad >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));           // lower element of the pair
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a < b) == desc) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace htcn
