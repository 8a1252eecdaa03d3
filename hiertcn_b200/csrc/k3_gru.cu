// K3: GRU over sessions (fp32).  One CTA owns a tile of kUB users and runs all S steps on chip:
// the recurrent state never leaves shared memory between steps; users are independent, so there is
// no grid-wide sync.  Weights (G x 384 KB fp32) do not fit one SM: they are streamed from L2 each
// step (coalesced 1 KB rows), which at 128 CTAs x 10 steps is ~1 GB of L2 reads per launch.
//
// Replaces model_hier.py:30-37,91,93 (MultiRNNCell of stock GRUCells + mask reset); formula per
// customed_gru_cell.py:309-337, layer stacking :1050-1073, linear :1187-1197.  Also emits the hoisted
// state half of the TCN in-projection, sbias[s] = state_pre[s] @ W_in[D:, :] (model_hier.py:54-55 +
// model_tcn.py:35), so K2 never sees the 256-wide state.
#include <cstdlib>

#include "common.cuh"

namespace htcn {

constexpr int kGruThreads = 256;

struct GruWeights {
  const float* gate_w[HTCN_MAX_GRU_LAYERS];   // [256, 256]
  const float* gate_b[HTCN_MAX_GRU_LAYERS];   // [256]
  const float* cand_w[HTCN_MAX_GRU_LAYERS];   // [256, 128]
  const float* cand_b[HTCN_MAX_GRU_LAYERS];   // [128]
  int num_layer;
};

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// acc[u] += sum_k src[u][k] * w[k*ldw + col]   for u in [u0, u0+NU), k in [0,128)
template <int NU>
__device__ __forceinline__ void dot_block(float (&acc)[NU], const float* __restrict__ src /*[users][128] smem*/,
                                          int u0, const float* __restrict__ w, int ldw, int col) {
#pragma unroll 2
  for (int k = 0; k < kDim; k += 4) {
    const float w0 = __ldg(w + (long long)(k + 0) * ldw + col);
    const float w1 = __ldg(w + (long long)(k + 1) * ldw + col);
    const float w2 = __ldg(w + (long long)(k + 2) * ldw + col);
    const float w3 = __ldg(w + (long long)(k + 3) * ldw + col);
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const float4 a = *reinterpret_cast<const float4*>(src + (u0 + u) * kDim + k);   // warp-broadcast
      acc[u] = fmaf(a.x, w0, acc[u]);
      acc[u] = fmaf(a.y, w1, acc[u]);
      acc[u] = fmaf(a.z, w2, acc[u]);
      acc[u] = fmaf(a.w, w3, acc[u]);
    }
  }
}

// kUB = users per CTA: 32 for large batches; 8 when the batch is small (data-parallel training at 512 users per GPU: 16 CTAs
// of 32 users leave 132 SMs idle for the whole recurrence -- 64 CTAs of 8 users run it 2.5x faster)
template <int kUB>
__global__ void __launch_bounds__(kGruThreads, kUB <= 16 ? 2 : 1)
k3_gru_sessions(const float* __restrict__ yp, const float* __restrict__ mask, const float* __restrict__ state_in,
                GruWeights W, const float* __restrict__ w_in_state, int B, int S,
                float* __restrict__ state_pre, float* __restrict__ sbias, float* __restrict__ state_out,
                float* __restrict__ gates_save /* [S][G][3][B][128] = r, u, c of every cell call, or NULL */) {
  extern __shared__ float smem[];
  const int G = W.num_layer;
  float* xin = smem;                         // [kUB][128]   layer-0 input of this step
  float* h = xin + kUB * kDim;               // [G][kUB][128]
  float* rh = h + G * kUB * kDim;            // [kUB][128]   r * h
  float* ug = rh + kUB * kDim;               // [kUB][128]   update gate
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * kUB;
  const int GH = G * kDim;

  for (int i = tid; i < G * kUB * kDim; i += kGruThreads) {
    const int g = i / (kUB * kDim), u = (i / kDim) % kUB, c = i % kDim;
    h[i] = (b0 + u < B) ? state_in[(long long)(b0 + u) * GH + g * kDim + c] : 0.f;
  }
  __syncthreads();

  for (int s = 0; s < S; ++s) {
    // ---- stage yp[s], emit state_pre[s] -----------------------------------------------------
    for (int i = tid; i < kUB * kDim; i += kGruThreads) {
      const int u = i / kDim, c = i % kDim;
      xin[i] = (b0 + u < B) ? yp[((long long)s * B + b0 + u) * kDim + c] : 0.f;
    }
    if (state_pre) {
      for (int i = tid; i < G * kUB * kDim; i += kGruThreads) {
        const int g = i / (kUB * kDim), u = (i / kDim) % kUB, c = i % kDim;
        if (b0 + u < B) state_pre[((long long)s * B + b0 + u) * GH + g * kDim + c] = h[i];
      }
    }
    // ---- sbias[s] = [h_0 | h_1 | ..] @ w_in_state  ([kUB, G*128] x [G*128, 128]) --------------
    if (sbias) {
      const int col = tid & 127, u0 = (tid >> 7) * (kUB / 2);
      float acc[kUB / 2];
#pragma unroll
      for (int u = 0; u < kUB / 2; ++u) acc[u] = 0.f;
      for (int g = 0; g < G; ++g)
        dot_block<kUB / 2>(acc, h + g * kUB * kDim, u0, w_in_state + (long long)g * kDim * kDim, kDim, col);
#pragma unroll
      for (int u = 0; u < kUB / 2; ++u)
        if (b0 + u0 + u < B) sbias[((long long)s * B + b0 + u0 + u) * kDim + col] = acc[u];
    }
    __syncthreads();

    // ---- layer stack ------------------------------------------------------------------------
    for (int g = 0; g < G; ++g) {
      const float* inp = (g == 0) ? xin : (h + (g - 1) * kUB * kDim);   // new state of the layer below
      float* hg = h + g * kUB * kDim;
      {  // gates: [inp | h] @ Wg + bg, 256 columns, one per thread, all kUB users
        float acc[kUB];
#pragma unroll
        for (int u = 0; u < kUB; ++u) acc[u] = 0.f;
        dot_block<kUB>(acc, inp, 0, W.gate_w[g], 2 * kDim, tid);
        dot_block<kUB>(acc, hg, 0, W.gate_w[g] + (long long)kDim * 2 * kDim, 2 * kDim, tid);
        const float bias = __ldg(W.gate_b[g] + tid);
        float* gsave = gates_save ? gates_save + ((long long)(s * G + g) * 3 + (tid < kDim ? 0 : 1)) * B * kDim : nullptr;
        if (tid < kDim) {
#pragma unroll
          for (int u = 0; u < kUB; ++u) {
            const float r = sigmoid_f(acc[u] + bias);
            rh[u * kDim + tid] = r * hg[u * kDim + tid];
            if (gsave && b0 + u < B) gsave[(long long)(b0 + u) * kDim + tid] = r;
          }
        } else {
#pragma unroll
          for (int u = 0; u < kUB; ++u) {
            const float z = sigmoid_f(acc[u] + bias);
            ug[u * kDim + tid - kDim] = z;
            if (gsave && b0 + u < B) gsave[(long long)(b0 + u) * kDim + tid - kDim] = z;
          }
        }
      }
      __syncthreads();
      {  // candidate: [inp | r*h] @ Wc + bc, 128 columns x 2 user halves
        const int col = tid & 127, u0 = (tid >> 7) * (kUB / 2);
        float acc[kUB / 2];
#pragma unroll
        for (int u = 0; u < kUB / 2; ++u) acc[u] = 0.f;
        dot_block<kUB / 2>(acc, inp, u0, W.cand_w[g], kDim, col);
        dot_block<kUB / 2>(acc, rh, u0, W.cand_w[g] + (long long)kDim * kDim, kDim, col);
        const float bias = __ldg(W.cand_b[g] + col);
        __syncthreads();   // every thread is done reading inp (== h[g-1]) and rh before h[g] changes
#pragma unroll
        for (int u = 0; u < kUB / 2; ++u) {
          const int i = (u0 + u) * kDim + col;
          const float c = tanhf(acc[u] + bias);
          const float uu = ug[i];
          hg[i] = uu * hg[i] + (1.0f - uu) * c;
          if (gates_save && b0 + u0 + u < B)
            gates_save[((long long)(s * G + g) * 3 + 2) * B * kDim + (long long)(b0 + u0 + u) * kDim + col] = c;
        }
      }
      __syncthreads();
    }
    // ---- user-boundary reset (model_hier.py:93) -----------------------------------------------
    for (int i = tid; i < G * kUB * kDim; i += kGruThreads) {
      const int u = (i / kDim) % kUB;
      const float m = (b0 + u < B) ? __ldg(mask + (long long)s * B + b0 + u) : 0.f;
      h[i] *= m;
    }
    __syncthreads();
  }
  for (int i = tid; i < G * kUB * kDim; i += kGruThreads) {
    const int g = i / (kUB * kDim), u = (i / kDim) % kUB, c = i % kDim;
    if (b0 + u < B) state_out[(long long)(b0 + u) * GH + g * kDim + c] = h[i];
  }
}

int32_t gru_sessions_bf16(const float* yp, const float* mask, const float* state_in, const float* const* gate_w,
                          const float* const* gate_b, const float* const* cand_w, const float* const* cand_b,
                          const float* w_in_state, int B, int S, float* state_pre, float* sbias, float* state_out,
                          float* scratch, cudaStream_t st, float* gates_save = nullptr);

}  // namespace htcn

static int32_t gru_sessions_impl(const float* yp, const float* mask, const float* state_in,
                                 const float* const* gate_w_host, const float* const* gate_b_host,
                                 const float* const* cand_w_host, const float* const* cand_b_host,
                                 int32_t num_layer, const float* w_in_state, int32_t B, int32_t S,
                                 int32_t precision, float* scratch,
                                 float* state_pre, float* sbias, float* state_out, float* gates_save, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(yp && mask && state_in && state_out && gate_w_host && gate_b_host && cand_w_host && cand_b_host,
               "gru_sessions: NULL pointer");
  HTCN_REQUIRE(num_layer >= 1 && num_layer <= HTCN_MAX_GRU_LAYERS, "gru_sessions: num_layer=%d", num_layer);
  HTCN_REQUIRE(B > 0 && S > 0, "gru_sessions: B=%d S=%d", B, S);
  HTCN_REQUIRE(!sbias || w_in_state, "gru_sessions: sbias requested without w_in_state");
  GruWeights W;
  W.num_layer = num_layer;
  for (int g = 0; g < num_layer; ++g) {
    W.gate_w[g] = gate_w_host[g];
    W.gate_b[g] = gate_b_host[g];
    W.cand_w[g] = cand_w_host[g];
    W.cand_b[g] = cand_b_host[g];
    HTCN_REQUIRE(W.gate_w[g] && W.gate_b[g] && W.cand_w[g] && W.cand_b[g], "gru_sessions: layer %d weights NULL", g);
  }
  HTCN_REQUIRE(precision == HTCN_F32 || precision == HTCN_BF16, "gru_sessions: precision %d", precision);
  if (precision == HTCN_BF16) {
    HTCN_REQUIRE(num_layer == 2, "gru_sessions(bf16): the tensor-core kernel is built for num_layer == 2 (got %d)", num_layer);
    HTCN_REQUIRE(scratch, "gru_sessions(bf16): scratch (HTCN_GRU_SCRATCH_BYTES(B)) is required");
    return gru_sessions_bf16(yp, mask, state_in, W.gate_w, W.gate_b, W.cand_w, W.cand_b, w_in_state, B, S, state_pre, sbias,
                             state_out, scratch, as_stream(stream), gates_save);
  }
  // users per CTA: HTCN_K3_F32_UB (8 / 16 / 32), default by batch size
  const char* ube = getenv("HTCN_K3_F32_UB");
  const int ub = ube ? atoi(ube) : (B <= 1024 ? 8 : B <= 2048 ? 16 : 32);
#define HTCN_K3_F32_LAUNCH(UB)                                                                                         \
  do {                                                                                                                \
    const size_t smem = sizeof(float) * (size_t)(3 + num_layer) * UB * kDim;                                          \
    HTCN_CUDA(cudaFuncSetAttribute(k3_gru_sessions<UB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    k3_gru_sessions<UB><<<ceil_div(B, UB), kGruThreads, smem, as_stream(stream)>>>(                                   \
        yp, mask, state_in, W, w_in_state, B, S, state_pre, sbias, state_out, gates_save);                            \
  } while (0)
  if (ub == 4) HTCN_K3_F32_LAUNCH(4);
  else if (ub == 8) HTCN_K3_F32_LAUNCH(8);
  else if (ub == 16) HTCN_K3_F32_LAUNCH(16);
  else HTCN_K3_F32_LAUNCH(32);
#undef HTCN_K3_F32_LAUNCH
  HTCN_LAUNCH_CHECK("k3_gru_sessions");
  return HTCN_OK;
}

extern "C" int32_t htcn_gru_sessions(const float* yp, const float* mask, const float* state_in,
                                     const float* const* gate_w_host, const float* const* gate_b_host,
                                     const float* const* cand_w_host, const float* const* cand_b_host,
                                     int32_t num_layer, const float* w_in_state, int32_t B, int32_t S,
                                     int32_t precision, float* scratch,
                                     float* state_pre, float* sbias, float* state_out, void* stream) {
  return gru_sessions_impl(yp, mask, state_in, gate_w_host, gate_b_host, cand_w_host, cand_b_host, num_layer, w_in_state,
                           B, S, precision, scratch, state_pre, sbias, state_out, nullptr, stream);
}

// fp32 forward that also saves the gate activations (r, u, c of every cell call) and the pre-step states for
// htcn_gru_backward
extern "C" int32_t htcn_gru_sessions_train(const float* yp, const float* mask, const float* state_in,
                                           const float* const* gate_w_host, const float* const* gate_b_host,
                                           const float* const* cand_w_host, const float* const* cand_b_host,
                                           int32_t num_layer, const float* w_in_state, int32_t B, int32_t S,
                                           float* state_pre, float* sbias, float* state_out, float* gates_save,
                                           void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(state_pre && gates_save, "gru_sessions_train: state_pre and gates_save are required");
  return gru_sessions_impl(yp, mask, state_in, gate_w_host, gate_b_host, cand_w_host, cand_b_host, num_layer, w_in_state,
                           B, S, HTCN_F32, nullptr, state_pre, sbias, state_out, gates_save, stream);
}

// The same on the tensor cores (bf16 operands, fp32 state and accumulation: k3_gru_t.cu with the gate activations written
// out).  num_layer == 2; scratch as htcn_gru_sessions' bf16 tier.
extern "C" int32_t htcn_gru_sessions_train_bf16(const float* yp, const float* mask, const float* state_in,
                                                const float* const* gate_w_host, const float* const* gate_b_host,
                                                const float* const* cand_w_host, const float* const* cand_b_host,
                                                int32_t num_layer, const float* w_in_state, int32_t B, int32_t S,
                                                float* scratch, float* state_pre, float* sbias, float* state_out,
                                                float* gates_save, void* stream) {
  using namespace htcn;
  HTCN_REQUIRE(state_pre && gates_save, "gru_sessions_train_bf16: state_pre and gates_save are required");
  return gru_sessions_impl(yp, mask, state_in, gate_w_host, gate_b_host, cand_w_host, cand_b_host, num_layer, w_in_state,
                           B, S, HTCN_BF16, scratch, state_pre, sbias, state_out, gates_save, stream);
}
