// K1: embedding gather (x path) + session mean-pool (y path).  HBM-bound.
// Replaces the one-hot x table GEMMs of model.py:59-61 + model_hier.py:50,83-85.
//
// One row of the table is up to 128 fp32 = 512 B = one warp x one 128-bit load per lane
// (ld.global.nc.L1::no_allocate.v4.f32); id 0 / out-of-range ids write zeros and issue no load.
// The table may be stored PACKED: `pitch4` float4 per row (emb_dim 100 -> 25 float4 = 400 B rows, 13 DRAM sectors instead
// of 16); lanes >= pitch4 issue no load and contribute the zero padding of the 128-wide output row.
// Algorithmic bytes per user-sequence: 2*T*4*emb_dim (rows) + 2*T*4 (ids) + T*128*e + S*512 (out).
#include "common.cuh"

namespace htcn {

constexpr int kRowsPerIter = 8;   // independent 512 B row loads in flight per warp

template <bool kBf16Out>
__global__ void __launch_bounds__(256)
k1_gather_rows(const float4* __restrict__ table, int pitch4, int item_num, const int* __restrict__ ids, long long n_rows,
               void* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r0 = warp * kRowsPerIter; r0 < n_rows; r0 += n_warps * kRowsPerIter) {
    // lanes 0..7 fetch the ids of the 8 rows, then broadcast
    int my_id = 0;
    if (lane < kRowsPerIter && r0 + lane < n_rows) my_id = __ldg(ids + r0 + lane);
    float4 v[kRowsPerIter];
#pragma unroll
    for (int i = 0; i < kRowsPerIter; ++i) {
      const int id = __shfl_sync(0xffffffffu, my_id, i);
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (id > 0 && id < item_num && lane < pitch4) v[i] = ldg_nc_f4(table + (long long)id * pitch4 + lane);
    }
#pragma unroll
    for (int i = 0; i < kRowsPerIter; ++i) {
      const long long r = r0 + i;
      if (r < n_rows) {
        if (kBf16Out) {
          uint2 p = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
          stg_na_u2(reinterpret_cast<uint2*>(out) + r * 32 + lane, p);
        } else {
          stg_na_f4(reinterpret_cast<float4*>(out) + r * 32 + lane, v[i]);
        }
      }
    }
  }
}

// one warp per (slot, user): sequential left-to-right accumulation (the order the oracle uses)
__global__ void __launch_bounds__(256)
k1_meanpool(const float4* __restrict__ table, int pitch4, const float4* __restrict__ bias, int item_num,
            const int* __restrict__ y_id, SlotTable slots, int B, int T, float4* __restrict__ yp) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (warp >= (long long)slots.n * B) return;
  const int s = (int)(warp / B), b = (int)(warp % B);
  const int p0 = slots.off[s], p1 = slots.off[s + 1];
  const int* ids = y_id + (long long)b * T;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int n = 0;
  for (int p = p0; p < p1; p += 4) {
    int id[4];
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      id[i] = (p + i < p1) ? __ldg(ids + p + i) : 0;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (id[i] > 0 && id[i] < item_num && lane < pitch4) v[i] = ldg_nc_f4(table + (long long)id[i] * pitch4 + lane);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (id[i] > 0) {          // sign(id) > 0 counts (model_hier.py:83); out-of-range rows are zero
        acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w;
        ++n;
      }
    }
  }
  const float fn = (float)n;    // n == 0 -> 0/0 = NaN, like the reference
  const float4 bb = __ldg(bias + lane);
  float4 o;
  o.x = __fdiv_rn(acc.x, fn) + bb.x;
  o.y = __fdiv_rn(acc.y, fn) + bb.y;
  o.z = __fdiv_rn(acc.z, fn) + bb.z;
  o.w = __fdiv_rn(acc.w, fn) + bb.w;
  yp[warp * 32 + lane] = o;
}

}  // namespace htcn

extern "C" int32_t htcn_gather_meanpool(const float* emb_table, int32_t emb_pitch, const float* emb_bias, int32_t item_num,
                                        const int32_t* x_id, const int32_t* y_id, const int32_t* slot_off,
                                        int32_t B, int32_t T, int32_t S, void* xe, int32_t xe_dtype,
                                        float* yp, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(emb_table && B > 0 && T > 0 && item_num > 0, "gather_meanpool: bad sizes/pointers");
  HTCN_REQUIRE(emb_pitch >= 4 && emb_pitch <= kDim && emb_pitch % 4 == 0,
               "gather_meanpool: emb_pitch=%d must be a multiple of 4 floats in [4, 128]", emb_pitch);
  const int pitch4 = emb_pitch / 4;
  HTCN_REQUIRE(xe_dtype == HTCN_F32 || xe_dtype == HTCN_BF16, "gather_meanpool: xe_dtype %d", xe_dtype);
  cudaStream_t st = as_stream(stream);
  if (xe) {
    HTCN_REQUIRE(x_id, "gather_meanpool: x_id is NULL");
    const long long rows = (long long)B * T;
    const int blocks = (int)std::min<long long>((rows + 8 * kRowsPerIter - 1) / (8 * kRowsPerIter), 148LL * 16);
    if (xe_dtype == HTCN_BF16)
      k1_gather_rows<true><<<blocks, 256, 0, st>>>((const float4*)emb_table, pitch4, item_num, x_id, rows, xe);
    else
      k1_gather_rows<false><<<blocks, 256, 0, st>>>((const float4*)emb_table, pitch4, item_num, x_id, rows, xe);
    HTCN_LAUNCH_CHECK("k1_gather_rows");
  }
  if (yp) {
    HTCN_REQUIRE(y_id && emb_bias && slot_off, "gather_meanpool: y_id/emb_bias/slot_off is NULL");
    HTCN_REQUIRE(S > 0 && S <= HTCN_MAX_SLOTS, "gather_meanpool: S=%d out of range", S);
    SlotTable slots;
    slots.n = S;
    for (int i = 0; i <= S; ++i) slots.off[i] = slot_off[i];   // host array, passed to the kernel by value
    HTCN_REQUIRE(slots.off[0] == 0 && slots.off[S] == T, "gather_meanpool: slot_off does not span T");
    const long long warps = (long long)S * B;
    k1_meanpool<<<ceil_div(warps, 8), 256, 0, st>>>((const float4*)emb_table, pitch4, (const float4*)emb_bias,
                                                    item_num, y_id, slots, B, T, (float4*)yp);
    HTCN_LAUNCH_CHECK("k1_meanpool");
  }
  return HTCN_OK;
}
