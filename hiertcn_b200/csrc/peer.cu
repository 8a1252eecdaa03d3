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
