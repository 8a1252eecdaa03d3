// Device-side batch assembly: the reference builds every batch on the host (data_loader.py:233-272:
// per-slot queues of sessions -> zero-padded x/y id matrices with y = x shifted by one -> numpy -> feed_dict) and times
// it separately as "load".  Here the interaction log lives in HBM and a batch is assembled by one small kernel from a
// per-slot schedule of session ids (the queue discipline of data_loader.py:170-231 replayed once on the host), followed
// by the compaction of the scored positions (row_of / y_rows / Q) that K2 and K4 consume.
//
// Every slot is padded to the fixed width L = max_activity_len instead of the batch's longest session (the reference's
// L_s): padded positions carry y = 0, are never scored, and a causal stack cannot see them from earlier positions, so
// losses / ranks / states are identical -- and no size has to travel back to the host except Q.
#include "common.cuh"

namespace htcn {
namespace {

// grid (B, S): one warp per (user slot, session slot)
__global__ void assemble_batch_kernel(const int* __restrict__ items, const int* __restrict__ sess_off,
                                      const int* __restrict__ sched_sess, const unsigned char* __restrict__ sched_last,
                                      int sched_pitch, int first, int B, int S, int L, int* __restrict__ x_id,
                                      int* __restrict__ y_id, float* __restrict__ mask) {
  const int b = blockIdx.x, s = blockIdx.y;
  const int k = first + s;
  const int sid = sched_sess[(long long)b * sched_pitch + k];
  const int start = sess_off[sid];
  int n = sess_off[sid + 1] - start;
  if (n > L) n = L;                                             // sessions are clipped to max_activity_len (:201-203)
  const int T = S * L;
  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    const long long o = (long long)b * T + s * L + t;
    y_id[o] = (t < n) ? items[start + t] : 0;                   // y = [a_1 .. a_n, 0 ..]
    x_id[o] = (t >= 1 && t - 1 < n) ? items[start + t - 1] : 0; // x = [0, a_1 .. a_n, 0 ..] cut to L (data_loader.py:246-261)
  }
  if (threadIdx.x == 0) mask[(long long)s * B + b] = sched_last[(long long)b * sched_pitch + k] ? 0.f : 1.f;
}

constexpr int kScanBlock = 1024;      // elements per block (256 threads x 4)

__global__ void valid_count_kernel(const int* __restrict__ y_id, long long n, int* __restrict__ block_cnt) {
  __shared__ int warp_sum[8];
  const long long base = (long long)blockIdx.x * kScanBlock + threadIdx.x * 4;
  int c = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) c += (base + e < n && y_id[base + e] > 0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += warp_sum[w];
    block_cnt[blockIdx.x] = t;
  }
}

// exclusive scan of the block counts in place (one block; n_blocks is a few thousand at most) + total
__global__ void block_scan_kernel(int* __restrict__ block_cnt, int n_blocks, int* __restrict__ total) {
  __shared__ int carry;
  __shared__ int warp_sum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n_blocks; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const int v = (i < n_blocks) ? block_cnt[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_sum[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      warp_sum[threadIdx.x] = w;
    }
    __syncthreads();
    const int warp_base = (threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0;
    const int incl = carry + warp_base + x;
    if (i < n_blocks) block_cnt[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void compact_kernel(const int* __restrict__ y_id, long long n, const int* __restrict__ block_off,
                               int* __restrict__ row_of, int* __restrict__ y_rows) {
  __shared__ int warp_sum[8];
  const long long base = (long long)blockIdx.x * kScanBlock + threadIdx.x * 4;
  int y[4], c = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    y[e] = (base + e < n) ? y_id[base + e] : 0;
    c += y[e] > 0;
  }
  int x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += v;
  }
  if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
  __syncthreads();
  int pre = block_off[blockIdx.x] + x - c;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) pre += warp_sum[w];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (base + e >= n) break;
    if (y[e] > 0) {
      row_of[base + e] = pre;
      y_rows[pre] = y[e];
      ++pre;
    } else {
      row_of[base + e] = -1;
    }
  }
}

}  // namespace
}  // namespace htcn

using namespace htcn;

extern "C" int64_t htcn_batcher_scratch_ints(int32_t B, int32_t T) {
  return ((int64_t)B * T + kScanBlock - 1) / kScanBlock + 4;
}

extern "C" int32_t htcn_assemble_batch(const int32_t* items, const int32_t* sess_off, const int32_t* sched_sess,
                                       const uint8_t* sched_last, int32_t sched_pitch, int32_t first_session, int32_t B,
                                       int32_t S, int32_t L, int32_t* x_id, int32_t* y_id, float* mask, int32_t* row_of,
                                       int32_t* y_rows, int32_t* n_valid, int32_t* scratch, void* stream) {
  HTCN_REQUIRE(items && sess_off && sched_sess && sched_last && x_id && y_id && mask && row_of && y_rows && n_valid && scratch,
               "assemble_batch: NULL pointer");
  HTCN_REQUIRE(B > 0 && S > 0 && S <= HTCN_MAX_SLOTS && L > 0 && first_session >= 0 && first_session + S <= sched_pitch,
               "assemble_batch: B=%d S=%d L=%d first_session=%d sched_pitch=%d", B, S, L, first_session, sched_pitch);
  cudaStream_t st = as_stream(stream);
  assemble_batch_kernel<<<dim3(B, S), 32, 0, st>>>(items, sess_off, sched_sess, sched_last, sched_pitch, first_session, B, S,
                                                  L, x_id, y_id, mask);
  HTCN_LAUNCH_CHECK("assemble_batch_kernel");
  const long long n = (long long)B * S * L;
  const int blocks = ceil_div(n, kScanBlock);
  valid_count_kernel<<<blocks, 256, 0, st>>>(y_id, n, scratch);
  HTCN_LAUNCH_CHECK("valid_count_kernel");
  block_scan_kernel<<<1, 1024, 0, st>>>(scratch, blocks, n_valid);
  HTCN_LAUNCH_CHECK("block_scan_kernel");
  compact_kernel<<<blocks, 256, 0, st>>>(y_id, n, scratch, row_of, y_rows);
  HTCN_LAUNCH_CHECK("compact_kernel");
  return HTCN_OK;
}
