// Exchanges of the catalog-sharded scoring path (BASELINE config 4) over NVLink peer memory -- no NCCL on the data path.
//
// Every rank owns one "symmetric" device buffer (same layout on every rank) and maps the buffers of all its peers into its
// own address space (CUDA IPC; the handles are exchanged once, through torch.distributed's object all-gather).  The three
// exchanges of ShardedCatalogScorer.score -- all-gather of the query rows, completion of the target-logit vector, all-to-all
// of the per-shard partials -- are then ONE kernel each: the kernel stores its rows straight into the peers' buffers with
// 128-bit writes over NVLink, publishes them with a system-scope release of an epoch flag in each peer's buffer, and waits
// (acquire) for the flags the peers set in its own buffer.  What NCCL does with a kernel per collective plus staging copies
// (5 all_to_all_single + 5 transposes for the partials) is one launch, and the data lands where the next kernel reads it.
//
// Ordering argument (why single buffering is safe): rank A can only finish call k after it received every peer's partials of
// call k, which a peer sends after ITS sweep of call k; so when A starts call k+1 and overwrites B's query rows, B's sweep of
// call k is over.  Flags carry the call's epoch, so a stale flag never satisfies a wait.
#include <cstring>

#include "common.cuh"

namespace htcn {

constexpr int kPeerThreads = 256;
constexpr long long kSpinLimit = 3000000000LL;     // ~1.5 s: a missing peer must fail the call, not hang the GPU

struct PeerSeg {                 // 2-D copy: n_rows rows of row_bytes (multiple of 16), row pitches in bytes
  const uint8_t* src;
  uint8_t* dst;
  unsigned long long row_bytes, src_pitch, dst_pitch;
  int n_rows;
};
struct PeerPushArgs {
  PeerSeg seg[HTCN_MAX_PEERS][HTCN_PEER_MAX_SEGS];
  uint32_t* flag_remote[HTCN_MAX_PEERS];   // &flags[kind][my rank] in peer p's buffer
  uint32_t* flag_local;                    // &flags[kind][0] in my buffer: entry p is set by rank p
  uint32_t* done;                          // [HTCN_MAX_PEERS] local block counters, zero between launches
  int* err;                                // set to 1 when a wait ran into the spin limit
  int n_peers, n_seg, rank;
  uint32_t epoch;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// block-level epilogue of every exchange kernel: the last block that finished its stores for peer p raises my flag there;
// block (0, p) then waits until rank p has raised its flag here
__device__ __forceinline__ void publish_and_wait(uint32_t* const* flag_remote, const uint32_t* flag_local, uint32_t* done, int* err,
                                                 int p_first, int p_last, int blocks_per_peer, bool waiter, uint32_t epoch) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int p = p_first; p <= p_last; ++p) {
      const unsigned prev = atomicAdd(&done[p], 1u);
      if (prev == (unsigned)blocks_per_peer - 1) {
        done[p] = 0;
        __threadfence_system();
        st_release_sys(flag_remote[p], epoch);
      }
    }
    if (waiter) {
      for (int p = p_first; p <= p_last; ++p) {
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(flag_local + p) - epoch) < 0) {
          if (clock64() - t0 > kSpinLimit) {
            atomicExch(err, 1);
            break;
          }
          __nanosleep(100);
        }
      }
    }
  }
}

// grid (blocks_per_peer, n_peers): block (b, p) copies its share of every segment destined to peer p
__global__ void __launch_bounds__(kPeerThreads) peer_push_kernel(const PeerPushArgs a) {
  const int p = blockIdx.y;
  for (int s = 0; s < a.n_seg; ++s) {
    const PeerSeg& g = a.seg[p][s];
    const unsigned long long per_row = g.row_bytes >> 4, total = per_row * (unsigned long long)g.n_rows;
    for (unsigned long long i = (unsigned long long)blockIdx.x * kPeerThreads + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * kPeerThreads) {
      const unsigned long long r = i / per_row, c = i - r * per_row;
      const uint4 v = *reinterpret_cast<const uint4*>(g.src + r * g.src_pitch + (c << 4));
      *reinterpret_cast<uint4*>(g.dst + r * g.dst_pitch + (c << 4)) = v;
    }
  }
  publish_and_wait(a.flag_remote, a.flag_local, a.done, a.err, p, p, gridDim.x, blockIdx.x == 0, a.epoch);
}

struct PeerBcastArgs {
  float* dst[HTCN_MAX_PEERS];              // the target-logit vector in every rank's buffer
  uint32_t* flag_remote[HTCN_MAX_PEERS];
  uint32_t* flag_local;
  uint32_t* done;
  int* err;
  int n_peers;
  uint32_t epoch;
};
// the owner of y[q] (n0 <= y[q] < n1) stores src[q] into every rank's vector; entries nobody owns keep their zero
__global__ void __launch_bounds__(kPeerThreads) peer_bcast_owned_kernel(const PeerBcastArgs a, const float* __restrict__ src,
                                                                        const int* __restrict__ y, int Q, int n0, int n1) {
  for (int q = blockIdx.x * kPeerThreads + threadIdx.x; q < Q; q += gridDim.x * kPeerThreads) {
    const int id = y[q];
    if (id >= n0 && id < n1) {
      const float v = src[q];
      for (int p = 0; p < a.n_peers; ++p) a.dst[p][q] = v;
    }
  }
  publish_and_wait(a.flag_remote, a.flag_local, a.done, a.err, 0, a.n_peers - 1, gridDim.x, blockIdx.x == 0, a.epoch);
}

// ---- data-parallel training: gradient all-reduce + scalar all-reduce + Adam in ONE kernel over peer memory --------------
struct PeerAdamArgs {
  const float4* grads[HTCN_MAX_PEERS];     // every rank's flat gradient buffer (mine included)
  float4* reduced[HTCN_MAX_PEERS];         // every rank's reduced-gradient buffer
  const float* scalars[HTCN_MAX_PEERS];    // every rank's {loss, r@1, r@5, r@10, mrr, mrp, user_count, n_valid}
  uint32_t* flag_in_remote[HTCN_MAX_PEERS];
  uint32_t* flag_mid_remote[HTCN_MAX_PEERS];
  uint32_t* flag_in_local;
  uint32_t* flag_mid_local;
  uint32_t* done;
  int* err;
  float4* p; float4* g; float4* m; float4* v;
  float* scalars_out;                      // [8] global means / counts (what allreduce_scalars returns)
  long long n4;
  int n_peers, rank;
  uint32_t epoch;
  float lr_t, b1, b2, eps;
};

__device__ __forceinline__ void wait_flags(const uint32_t* flag_local, int n_peers, uint32_t epoch, int* err) {
  for (int p = 0; p < n_peers; ++p) {
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(flag_local + p) - epoch) < 0) {
      if (clock64() - t0 > kSpinLimit) {
        atomicExch(err, 1);
        break;
      }
      __nanosleep(100);
    }
  }
}

// Persistent grid (every block co-resident).  Phase A: rank r sums slice r of the flat gradient over all ranks (peer loads,
// fixed rank order: every replica receives the SAME bits) and stores the sums into every rank's `reduced` buffer -- a
// reduce-scatter and an all-gather without a staging copy.  Phase B: the ordinary TF-Adam update (bwd_k1_adam.cu) on the
// reduced gradient, and the rank's own gradient buffer is cleared for the next step.
__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_adam_kernel(const PeerAdamArgs a) {
  __shared__ float s_div;
  const int W = a.n_peers;
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {                 // my gradients and scalars are complete (stream order): tell everybody
      __threadfence_system();
      for (int p = 0; p < W; ++p) st_release_sys(a.flag_in_remote[p], a.epoch);
    }
    wait_flags(a.flag_in_local, W, a.epoch, a.err);
    float total = 0.f;                     // global user count = the divisor of the summed gradient (model.py:116-117)
    for (int p = 0; p < W; ++p) total += a.scalars[p][6];
    s_div = total;
    if (blockIdx.x == 0) {
      for (int j = 0; j < 6; ++j) {
        float acc = 0.f;
        for (int p = 0; p < W; ++p) {
          const float c = a.scalars[p][6];
          if (c > 0.f) acc += a.scalars[p][j] * c;        // a rank without a scored user reports NaN means: it contributes nothing
        }
        a.scalars_out[j] = acc / total;
      }
      float nv = 0.f;
      for (int p = 0; p < W; ++p) nv += a.scalars[p][7];
      a.scalars_out[6] = total;
      a.scalars_out[7] = nv;
    }
  }
  __syncthreads();
  // ---- phase A: my slice
  const long long per = (a.n4 + W - 1) / W, lo = per * a.rank, hi = lo + per < a.n4 ? lo + per : a.n4;
  for (long long i = lo + (long long)blockIdx.x * kPeerThreads + threadIdx.x; i < hi; i += (long long)gridDim.x * kPeerThreads) {
    float4 part[HTCN_MAX_PEERS];
#pragma unroll
    for (int p = 0; p < HTCN_MAX_PEERS; ++p)
      if (p < W) part[p] = a.grads[p][i];
    float4 s = part[0];
#pragma unroll
    for (int p = 1; p < HTCN_MAX_PEERS; ++p)
      if (p < W) { s.x += part[p].x; s.y += part[p].y; s.z += part[p].z; s.w += part[p].w; }
#pragma unroll
    for (int p = 0; p < HTCN_MAX_PEERS; ++p)
      if (p < W) a.reduced[p][i] = s;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(&a.done[0], 1u);
    if (prev == gridDim.x - 1) {
      a.done[0] = 0;
      __threadfence_system();
      for (int p = 0; p < W; ++p) st_release_sys(a.flag_mid_remote[p], a.epoch);
    }
    wait_flags(a.flag_mid_local, W, a.epoch, a.err);      // every slice of `reduced` has landed; nobody reads my gradients any more
  }
  __syncthreads();
  // ---- phase B: Adam on the reduced gradient
  const float div = s_div;
  const float4* red = a.reduced[a.rank];
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = (long long)blockIdx.x * kPeerThreads + threadIdx.x; i < a.n4; i += (long long)gridDim.x * kPeerThreads) {
    if (div > 0.f) {                      // no scored user anywhere: no loss, no update (see adam_kernel)
      const float sc = 1.0f / div;
      float4 gg = red[i], mm = a.m[i], vv = a.v[i], pp = a.p[i];
#define HTCN_ADAM1(c)                                   \
  {                                                     \
    const float gr = gg.c * sc;                         \
    mm.c = a.b1 * mm.c + (1.f - a.b1) * gr;             \
    vv.c = a.b2 * vv.c + (1.f - a.b2) * gr * gr;        \
    pp.c -= a.lr_t * mm.c / (sqrtf(vv.c) + a.eps);      \
  }
      HTCN_ADAM1(x) HTCN_ADAM1(y) HTCN_ADAM1(z) HTCN_ADAM1(w)
#undef HTCN_ADAM1
      a.p[i] = pp; a.m[i] = mm; a.v[i] = vv;
    }
    a.g[i] = zero;
  }
}

}  // namespace htcn

using namespace htcn;

extern "C" int32_t htcn_peer_alloc(int64_t bytes, void** ptr) {
  HTCN_REQUIRE(ptr && bytes > 0, "peer_alloc: bad args");
  HTCN_CUDA(cudaMalloc(ptr, (size_t)bytes));
  HTCN_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
  HTCN_CUDA(cudaDeviceSynchronize());
  return HTCN_OK;
}

extern "C" int32_t htcn_peer_free(void* ptr) {
  if (ptr) HTCN_CUDA(cudaFree(ptr));
  return HTCN_OK;
}

extern "C" int32_t htcn_peer_export(const void* ptr, uint8_t* handle) {
  HTCN_REQUIRE(ptr && handle, "peer_export: bad args");
  static_assert(sizeof(cudaIpcMemHandle_t) == HTCN_PEER_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h;
  HTCN_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle, &h, sizeof(h));
  return HTCN_OK;
}

extern "C" int32_t htcn_peer_import(const uint8_t* handle, void** ptr) {
  HTCN_REQUIRE(ptr && handle, "peer_import: bad args");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  HTCN_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return HTCN_OK;
}

extern "C" int32_t htcn_peer_unimport(void* ptr) {
  if (ptr) HTCN_CUDA(cudaIpcCloseMemHandle(ptr));
  return HTCN_OK;
}

extern "C" int32_t htcn_peer_exchange(const void* const* src, void* const* dst, const int64_t* row_bytes, const int32_t* n_rows,
                                      const int64_t* src_pitch, const int64_t* dst_pitch, int32_t n_seg, int32_t n_peers,
                                      int32_t rank, void* const* flag_remote, void* flag_local, void* done, void* err,
                                      uint32_t epoch, void* stream) {
  HTCN_REQUIRE(src && dst && row_bytes && n_rows && src_pitch && dst_pitch && flag_remote && flag_local && done && err,
               "peer_exchange: bad args");
  HTCN_REQUIRE(n_peers >= 1 && n_peers <= HTCN_MAX_PEERS && n_seg >= 1 && n_seg <= HTCN_PEER_MAX_SEGS && rank >= 0 &&
                   rank < n_peers,
               "peer_exchange: n_peers=%d n_seg=%d rank=%d", n_peers, n_seg, rank);
  PeerPushArgs a;
  memset(&a, 0, sizeof(a));
  unsigned long long most = 0;
  for (int p = 0; p < n_peers; ++p) {
    for (int s = 0; s < n_seg; ++s) {
      const int i = p * n_seg + s;
      HTCN_REQUIRE(row_bytes[i] % 16 == 0 && src_pitch[i] % 16 == 0 && dst_pitch[i] % 16 == 0 &&
                       (reinterpret_cast<uintptr_t>(src[i]) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst[i]) & 15) == 0,
                   "peer_exchange: segment %d of peer %d is not 16-byte aligned", s, p);
      a.seg[p][s] = PeerSeg{static_cast<const uint8_t*>(src[i]), static_cast<uint8_t*>(dst[i]), (unsigned long long)row_bytes[i],
                            (unsigned long long)src_pitch[i], (unsigned long long)dst_pitch[i], n_rows[i]};
      most = most > (unsigned long long)row_bytes[i] * n_rows[i] ? most : (unsigned long long)row_bytes[i] * n_rows[i];
    }
    a.flag_remote[p] = static_cast<uint32_t*>(flag_remote[p]);
  }
  a.flag_local = static_cast<uint32_t*>(flag_local);
  a.done = static_cast<uint32_t*>(done);
  a.err = static_cast<int*>(err);
  a.n_peers = n_peers;
  a.n_seg = n_seg;
  a.rank = rank;
  a.epoch = epoch;
  // enough blocks per peer to keep the NVLink stores in flight, few enough that all of them are co-resident (the waiter
  // blocks spin): 16 KB of the largest segment per block, 1..16 blocks per peer
  int bpp = (int)((most + 16383) / 16384);
  bpp = bpp < 1 ? 1 : bpp > 16 ? 16 : bpp;
  peer_push_kernel<<<dim3(bpp, n_peers), kPeerThreads, 0, as_stream(stream)>>>(a);
  HTCN_LAUNCH_CHECK("peer_push_kernel");
  return HTCN_OK;
}

extern "C" int32_t htcn_peer_bcast_owned(const float* src, const int32_t* y_id, int32_t Q, int32_t n0, int32_t n1,
                                         void* const* dst, int32_t n_peers, void* const* flag_remote, void* flag_local,
                                         void* done, void* err, uint32_t epoch, void* stream) {
  HTCN_REQUIRE(src && y_id && dst && flag_remote && flag_local && done && err && Q > 0, "peer_bcast_owned: bad args");
  HTCN_REQUIRE(n_peers >= 1 && n_peers <= HTCN_MAX_PEERS, "peer_bcast_owned: n_peers=%d", n_peers);
  PeerBcastArgs a;
  memset(&a, 0, sizeof(a));
  for (int p = 0; p < n_peers; ++p) {
    a.dst[p] = static_cast<float*>(dst[p]);
    a.flag_remote[p] = static_cast<uint32_t*>(flag_remote[p]);
  }
  a.flag_local = static_cast<uint32_t*>(flag_local);
  a.done = static_cast<uint32_t*>(done);
  a.err = static_cast<int*>(err);
  a.n_peers = n_peers;
  a.epoch = epoch;
  int blocks = ceil_div(Q, kPeerThreads);
  blocks = blocks > 16 ? 16 : blocks;
  peer_bcast_owned_kernel<<<blocks, kPeerThreads, 0, as_stream(stream)>>>(a, src, y_id, Q, n0, n1);
  HTCN_LAUNCH_CHECK("peer_bcast_owned_kernel");
  return HTCN_OK;
}

extern "C" int32_t htcn_peer_allreduce_adam(void* const* grads, void* const* reduced, void* const* scalars, int32_t n_peers,
                                            int32_t rank, float* param, float* m, float* v, int64_t n, float lr_t, float beta1,
                                            float beta2, float eps, float* scalars_out, void* const* flag_in_remote,
                                            void* const* flag_mid_remote, void* flag_in_local, void* flag_mid_local, void* done,
                                            void* err, uint32_t epoch, void* stream) {
  HTCN_REQUIRE(grads && reduced && scalars && param && m && v && scalars_out && flag_in_remote && flag_mid_remote &&
                   flag_in_local && flag_mid_local && done && err && n > 0 && n % 4 == 0,
               "peer_allreduce_adam: bad args (n=%lld must be a multiple of 4)", (long long)n);
  HTCN_REQUIRE(n_peers >= 1 && n_peers <= HTCN_MAX_PEERS && rank >= 0 && rank < n_peers, "peer_allreduce_adam: n_peers=%d rank=%d",
               n_peers, rank);
  PeerAdamArgs a;
  memset(&a, 0, sizeof(a));
  for (int p = 0; p < n_peers; ++p) {
    a.grads[p] = static_cast<const float4*>(grads[p]);
    a.reduced[p] = static_cast<float4*>(reduced[p]);
    a.scalars[p] = static_cast<const float*>(scalars[p]);
    a.flag_in_remote[p] = static_cast<uint32_t*>(flag_in_remote[p]);
    a.flag_mid_remote[p] = static_cast<uint32_t*>(flag_mid_remote[p]);
  }
  a.flag_in_local = static_cast<uint32_t*>(flag_in_local);
  a.flag_mid_local = static_cast<uint32_t*>(flag_mid_local);
  a.done = static_cast<uint32_t*>(done);
  a.err = static_cast<int*>(err);
  a.p = reinterpret_cast<float4*>(param);
  a.g = const_cast<float4*>(a.grads[rank]);
  a.m = reinterpret_cast<float4*>(m);
  a.v = reinterpret_cast<float4*>(v);
  a.scalars_out = scalars_out;
  a.n4 = n / 4;
  a.n_peers = n_peers;
  a.rank = rank;
  a.epoch = epoch;
  a.lr_t = lr_t; a.b1 = beta1; a.b2 = beta2; a.eps = eps;
  // persistent: every block must be resident at once (the blocks meet at two flag waits); 2 blocks of 256 threads per SM
  int dev = 0, sms = 0;
  HTCN_CUDA(cudaGetDevice(&dev));
  HTCN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  peer_allreduce_adam_kernel<<<2 * sms, kPeerThreads, 0, as_stream(stream)>>>(a);
  HTCN_LAUNCH_CHECK("peer_allreduce_adam_kernel");
  return HTCN_OK;
}
