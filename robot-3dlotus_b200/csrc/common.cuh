// Shared helpers for the lotus sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define LOTUS_API extern "C" __attribute__((visibility("default")))

// ---- error channel (thread-local message, retrievable via lotus_last_error()) ----
void lotus_set_error(const char* fmt, ...);

#define LOTUS_CHECK_ARG(cond, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      lotus_set_error(__VA_ARGS__);           \
      return 2;                               \
    }                                         \
  } while (0)

#define LOTUS_CUDA_OK()                                                          \
  do {                                                                           \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) {                                                     \
      lotus_set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

#define LOTUS_CUDA_CALL(x)                                                       \
  do {                                                                           \
    cudaError_t _e = (x);                                                        \
    if (_e != cudaSuccess) {                                                     \
      lotus_set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static const int kNumSM = 148;  // B200

// ---- programmatic dependent launch (PDL) ----
// The step is ~700 dependent launches of 5-40 us kernels on one stream; with plain launches every boundary pays the previous grid's
// drain + the next grid's launch and prologue in series.  Kernels launched through launch_pdl() may be scheduled while the previous
// kernel of the stream is still draining; they call pdl_wait() BEFORE their first access to global memory (it returns once the
// previous grid has completed and its writes are visible) and pdl_launch_dependents() right after, which lets the next kernel's CTAs
// take SM slots as this grid's CTAs retire.  Both instructions are no-ops in a kernel launched without the attribute.
// LOTUS_NO_PDL=1 turns the attribute off (A/B knob, read once).
#include <stdlib.h>
static inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LOTUS_NO_PDL"); v = (e && atoi(e) != 0) ? 0 : 1; }
  return v != 0;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- device helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; `red` must hold >= 32 floats. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    r = warp_sum(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
  if (w == 0) {
    r = warp_max(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  return red[0];
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// activation element access: kernels are templated on the storage type (float | __nv_bfloat16), math is fp32
typedef __nv_bfloat16 bf16_t;
__device__ __forceinline__ float ldf(const float* p, int64_t i) { return p[i]; }
__device__ __forceinline__ float ldf(const bf16_t* p, int64_t i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void stf(float* p, int64_t i, float v) { p[i] = v; }
__device__ __forceinline__ void stf(bf16_t* p, int64_t i, float v) { p[i] = __float2bfloat16(v); }

// counter-based RNG for dropout masks: one 32-bit hash per element (seed, stream id, index) -> uniform [0,1)
__device__ __forceinline__ uint32_t lotus_hash3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return h;
}
// elementwise dropout (GEMM epilogues, pointwise, dropout kernels): one hash decides FOUR consecutive elements (a byte
// each), so the mask costs a quarter hash per element.  thr = round(p*256); keep-probability = 1 - thr/256.
__device__ __forceinline__ uint32_t drop_thr(float p) { return (uint32_t)fminf(255.f, floorf(p * 256.f + 0.5f)); }
__device__ __forceinline__ float drop_scale(uint32_t thr) { return 256.f / (256.f - (float)thr); }
__device__ __forceinline__ uint32_t drop_hash4(uint64_t seed, uint32_t sid, uint64_t idx) {   // covers elements idx&~3 .. +3
  return lotus_hash3((uint32_t)seed ^ (sid ^ (uint32_t)(idx >> 34)) * 0x27D4EB2Fu, (uint32_t)(seed >> 32), (uint32_t)(idx >> 2));
}
__device__ __forceinline__ bool drop_keep_e(uint64_t seed, uint32_t sid, uint64_t idx, uint32_t thr) {
  return ((drop_hash4(seed, sid, idx) >> (8 * (uint32_t)(idx & 3))) & 0xFFu) >= thr;
}
// patch-attention probability dropout: ONE full hash per (patch, head, query) and a cheap multiplicative mix per key, so the
// element loops of the attention kernels (128 x 128 probabilities per patch and head, forward and backward) spend ~6 integer
// instructions per mask bit instead of ~14.  keep(q, key) is the same function in the SIMT and the tensor-core kernels.
__device__ __forceinline__ uint32_t attn_drop_rowhash(uint64_t seed, uint32_t sid, uint32_t q) {
  return lotus_hash3((uint32_t)seed ^ sid * 0x27D4EB2Fu, (uint32_t)(seed >> 32), q);
}
__device__ __forceinline__ uint32_t attn_drop_thr(float p) { return (uint32_t)(p * 16777216.f); }
__device__ __forceinline__ bool attn_drop_keep(uint32_t hq, uint32_t key, uint32_t thr24) {
  uint32_t y = (hq + key * 0x9E3779B1u) * 0x85EBCA77u;
  y ^= y >> 15;
  y *= 0xC2B2AE3Du;
  return (y >> 8) >= thr24;
}
// cross-attention probability dropout keeps the per-element hash
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint32_t idx_hi, uint32_t idx_lo, float p) {
  uint32_t h = lotus_hash3((uint32_t)seed ^ idx_hi * 0x27D4EB2Fu, (uint32_t)(seed >> 32), idx_lo);
  return (h >> 8) * (1.0f / 16777216.0f) >= p;
}
