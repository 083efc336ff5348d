// Row/column normalisations and pointwise ops of the path, fp32 (SURVEY §2.3 last row):
//   LayerNorm fwd/bwd  (block norms, CPE norm, per-head q/k norm eps 1e-6: model.py:393-394,624-645)
//   BatchNorm1d(eps 1e-3, momentum 0.01) fwd/bwd fused with GELU (stem / pool / unpool: model.py:932, 785-788)
//   GELU / LeakyReLU / dropout / column sums (bias gradients) / small elementwise helpers
// All HBM-bound: one pass per tensor where possible, warp-per-row or column-tiled reductions.
#include "common.cuh"

struct RowMap {  // row r -> element offset (r / inner) * outer_stride + (r % inner) * inner_stride
  int inner;
  int64_t outer_stride, inner_stride;
  __device__ __forceinline__ int64_t operator()(int64_t r) const {
    return (r / inner) * outer_stride + (r % inner) * inner_stride;
  }
};

// Blocks of `kern` that are resident at once on the whole GPU (148 SMs x occupancy).  The persistent kernels below are launched with
// at most this many blocks: a grid of 4 blocks per SM whose occupancy is 3 runs as two waves, and for these 10-20 us kernels the second
// wave doubled the time.  Cached per kernel (the dynamic shared memory they use is a few KB and never the limiter).
#include <unordered_map>
template <typename K>
static int resident_blocks(K kern, int threads, size_t smem) {
  thread_local std::unordered_map<const void*, int> cache;
  auto it = cache.find((const void*)kern);
  if (it != cache.end()) return it->second;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) occ = 1;
  cache[(const void*)kern] = occ * kNumSM;
  return occ * kNumSM;
}

// ---------------------------------------------------------------- LayerNorm
// y = [residual +] (x - mean) * rstd * gamma + beta ; one warp per row
template <typename T>
__global__ void k_ln_fwd(const T* __restrict__ x, RowMap xm, T* __restrict__ y, RowMap ym,
                         const T* __restrict__ residual, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float* __restrict__ mean, float* __restrict__ rstd,
                         int64_t rows, int C, float eps) {
  int lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const T* xr = x + xm(r);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += ldf(xr, c);
  float mu = warp_sum(s) / C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { float d = ldf(xr, c) - mu; v += d * d; }
  float rs = rsqrtf(warp_sum(v) / C + eps);
  if (lane == 0) { if (mean) mean[r] = mu; if (rstd) rstd[r] = rs; }
  T* yr = y + ym(r);
  const T* rr = residual ? residual + ym(r) : nullptr;
  for (int c = lane; c < C; c += 32) {
    float o = (ldf(xr, c) - mu) * rs * gamma[c] + beta[c];
    stf(yr, c, rr ? ldf(rr, c) + o : o);
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma ; dgamma += dy*xhat ; dbeta += dy
// `accumulate_dx`: dx += (used when the normalised tensor also feeds a residual branch)
template <int CPL, typename TDY, typename T>  // channels per lane upper bound (C <= 32*CPL); dy type, x/dx type
__global__ void k_ln_bwd(const TDY* __restrict__ dy, RowMap dym, const T* __restrict__ x, RowMap xm,
                         T* __restrict__ dx, RowMap dxm, const float* __restrict__ gamma,
                         const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dgamma,
                         float* __restrict__ dbeta, int64_t rows, int C, int rows_per_block, int accumulate_dx) {
  extern __shared__ float sacc[];  // [2][C]
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) sacc[c] = 0.f;
  __syncthreads();
  float ag[CPL], ab[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
  int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = min(rows, r0 + rows_per_block);
  for (int64_t r = r0 + w; r < r1; r += nw) {
    const TDY* dyr = dy + dym(r);
    const T* xr = x + xm(r);
    float mu = mean[r], rs = rstd[r];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      int c = lane + 32 * i;
      if (c < C) {
        float xh = (ldf(xr, c) - mu) * rs, d = ldf(dyr, c), g = d * gamma[c];
        s1 += g; s2 += g * xh;
        ag[i] += d * xh; ab[i] += d;
      }
    }
    s1 = warp_sum(s1) / C; s2 = warp_sum(s2) / C;
    T* dxr = dx + dxm(r);
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      int c = lane + 32 * i;
      if (c < C) {
        float xh = (ldf(xr, c) - mu) * rs;
        float o = rs * (ldf(dyr, c) * gamma[c] - s1 - xh * s2);
        stf(dxr, c, accumulate_dx ? ldf(dxr, c) + o : o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    int c = lane + 32 * i;
    if (c < C) { atomicAdd(&sacc[c], ag[i]); atomicAdd(&sacc[C + c], ab[i]); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (dgamma) atomicAdd(&dgamma[c], sacc[c]);
    if (dbeta) atomicAdd(&dbeta[c], sacc[C + c]);
  }
}


// ---------------------------------------------------------------- vectorised bf16 LayerNorm (16-byte accesses, one pass)
// A row of C channels is covered by LPR lanes (power of two, <= 32), each holding VPL vectors of 8 bf16 in registers;
// 32/LPR rows per warp; mean / variance by shuffle inside the lane group.  Rows may be strided (RowMap), row bases and
// strides are multiples of 8 elements so every access is a 16-byte vector.
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = (const __nv_bfloat162*)&u;
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = (__nv_bfloat162*)&u;
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ void load8(const bf16_t* p, float* f) { unpack8(*(const uint4*)p, f); }
__device__ __forceinline__ void load8(const float* p, float* f) {
  float4 a = *(const float4*)p, b = *(const float4*)(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ float group_sum(float v, int lpr) {
  for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int VPL>
__global__ void __launch_bounds__(256) k_ln_fwd_vec(const bf16_t* __restrict__ x, RowMap xm, bf16_t* __restrict__ y, RowMap ym,
                                                    const bf16_t* __restrict__ residual, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, float* __restrict__ mean,
                                                    float* __restrict__ rstd, int64_t rows, int C, int lpr, float eps) {
  pdl_wait();                                                    // programmatic dependent launch (common.cuh)
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, lig = lane & (lpr - 1), rpw = 32 / lpr;
  const int nvec = C >> 3;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float invC = 1.f / (float)C;
  for (int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * rpw; r0 < rows; r0 += warps * rpw) {
    int64_t r = r0 + lane / lpr;
    bool rok = r < rows;
    float v[VPL][8];
    float s = 0.f;
    const bf16_t* xr = x + (rok ? xm(r) : 0);
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      int vi = lig + j * lpr;
      if (rok && vi < nvec) { load8(xr + vi * 8, v[j]); } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[j][e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[j][e];
    }
    float mu = group_sum(s, lpr) * invC;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      int vi = lig + j * lpr;
      if (vi < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { float d = v[j][e] - mu; q += d * d; }
      }
    }
    float rs = rsqrtf(group_sum(q, lpr) * invC + eps);
    if (!rok) continue;
    if (lig == 0) { if (mean) mean[r] = mu; if (rstd) rstd[r] = rs; }
    int64_t yo = ym(r);
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      int vi = lig + j * lpr;
      if (vi < nvec) {
        float g[8], b[8], o[8];
        load8(gamma + vi * 8, g); load8(beta + vi * 8, b);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (v[j][e] - mu) * rs * g[e] + b[e];
        if (residual) {
          float rr[8];
          load8(residual + yo + vi * 8, rr);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] += rr[e];
        }
        *(uint4*)(y + yo + vi * 8) = pack8(o);
      }
    }
  }
}

// U row groups per warp iteration: their loads are issued together (one group in flight per warp leaves the memory system half idle)
template <int VPL, int U, typename TDY>
__global__ void __launch_bounds__(256, U == 2 ? 3 : 1) k_ln_bwd_vec(const TDY* __restrict__ dy, RowMap dym, const bf16_t* __restrict__ x, RowMap xm,
                                                    bf16_t* __restrict__ dx, RowMap dxm, const float* __restrict__ gamma,
                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                    float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int C,
                                                    int lpr, int accumulate_dx, bf16_t* __restrict__ dx_masked, float drop_p, uint64_t seed,
                                                    uint32_t drop_sid) {
  extern __shared__ float sacc[];  // [2][C]
  const int lane = threadIdx.x & 31, lig = lane & (lpr - 1), rpw = 32 / lpr;
  const int nvec = C >> 3;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) sacc[c] = 0.f;
  __syncthreads();
  pdl_wait();                                                    // programmatic dependent launch (common.cuh)
  pdl_launch_dependents();
  const uint32_t dthr = drop_thr(drop_p);
  const float dks = dx_masked ? drop_scale(dthr) : 1.f;
  float ag[VPL][8], ab[VPL][8], g[VPL][8];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    int vi = lig + j * lpr;
#pragma unroll
    for (int e = 0; e < 8; ++e) { ag[j][e] = 0.f; ab[j][e] = 0.f; g[j][e] = 0.f; }
    if (vi < nvec) load8(gamma + vi * 8, g[j]);
  }
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float invC = 1.f / (float)C;
  for (int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * rpw; r0 < rows; r0 += warps * rpw * U) {
    int64_t r[U];
    bool rok[U];
    float mu[U], rs[U];
    uint4 xv[U][VPL];
    float d[U][VPL][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      r[u] = r0 + (int64_t)u * warps * rpw + lane / lpr;
      rok[u] = r[u] < rows;
      mu[u] = rok[u] ? mean[r[u]] : 0.f; rs[u] = rok[u] ? rstd[r[u]] : 0.f;
      const bf16_t* xr = x + (rok[u] ? xm(r[u]) : 0);
      const TDY* dyr = dy + (rok[u] ? dym(r[u]) : 0);
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        int vi = lig + j * lpr;
        if (rok[u] && vi < nvec) { xv[u][j] = *(const uint4*)(xr + vi * 8); load8(dyr + vi * 8, d[u][j]); }
        else {
          xv[u][j] = make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int e = 0; e < 8; ++e) d[u][j][e] = 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float xh[VPL][8];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        int vi = lig + j * lpr;
        unpack8(xv[u][j], xh[j]);
        const bool on = rok[u] && vi < nvec;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          xh[j][e] = on ? (xh[j][e] - mu[u]) * rs[u] : 0.f;
          float gg = d[u][j][e] * g[j][e];
          s1 += gg; s2 += gg * xh[j][e];
          ag[j][e] += d[u][j][e] * xh[j][e]; ab[j][e] += d[u][j][e];
        }
      }
      s1 = group_sum(s1, lpr) * invC; s2 = group_sum(s2, lpr) * invC;
      if (!rok[u]) continue;
      int64_t xo = dxm(r[u]);
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        int vi = lig + j * lpr;
        if (vi < nvec) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = rs[u] * (d[u][j][e] * g[j][e] - s1 - xh[j][e] * s2);
          if (accumulate_dx) {
            float old[8];
            load8(dx + xo + vi * 8, old);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] += old[e];
          }
          const uint4 packed = pack8(o);
          *(uint4*)(dx + xo + vi * 8) = packed;
          if (dx_masked) {
            // the gradient that continues into the NEXT layer's output dropout: mask(dx), contiguous [rows, C], same element index,
            // hash and rounding as lotus_dropout_bf16 on the stored bf16 values (bit-identical to the separate launch it replaces)
            float m[8];
            unpack8(packed, m);
            const int64_t fo = r[u] * C + vi * 8;
            const uint32_t h0 = drop_hash4(seed, drop_sid, (uint64_t)fo), h1 = drop_hash4(seed, drop_sid, (uint64_t)fo + 4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              m[e] = (((h0 >> (8 * e)) & 0xFFu) >= dthr) ? m[e] * dks : 0.f;
              m[4 + e] = (((h1 >> (8 * e)) & 0xFFu) >= dthr) ? m[4 + e] * dks : 0.f;
            }
            *(uint4*)(dx_masked + fo) = pack8(m);
          }
        }
      }
    }
  }
  // reduce the per-lane parameter-gradient partials: lanes with the same `lig` hold the same channels
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    int vi = lig + j * lpr;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float a = ag[j][e], b = ab[j][e];
      for (int o = 16; o >= lpr; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
      if (lane < lpr && vi < nvec) { atomicAdd(&sacc[vi * 8 + e], a); atomicAdd(&sacc[C + vi * 8 + e], b); }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (dgamma) atomicAdd(&dgamma[c], sacc[c]);
    if (dbeta) atomicAdd(&dbeta[c], sacc[C + c]);
  }
}

static inline bool vec_ok(const RowMap& m) { return (m.outer_stride % 8 == 0) && (m.inner_stride % 8 == 0); }
static inline int pick_lpr(int C) { int nv = C >> 3, l = 1; while (l < nv && l < 32) l <<= 1; return l; }
static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <typename T>
static int ln_fwd_impl(const void* x, int x_inner, int64_t x_outer_stride, int64_t x_inner_stride, void* y, int y_inner,
                       int64_t y_outer_stride, int64_t y_inner_stride, const void* residual, const float* gamma,
                       const float* beta, float* mean, float* rstd, int64_t rows, int C, float eps, void* stream_) {
  if (rows <= 0) return 0;
  RowMap xm{x_inner, x_outer_stride, x_inner_stride}, ym{y_inner, y_outer_stride, y_inner_stride};
  if (sizeof(T) == 2 && C % 8 == 0 && C <= 1024 && vec_ok(xm) && vec_ok(ym) && aligned16(x) && aligned16(y) &&
      (!residual || aligned16(residual)) && aligned16(gamma) && aligned16(beta)) {
    int lpr = pick_lpr(C), vpl = ceil_div(C >> 3, lpr), rpw = 32 / lpr;
    cudaStream_t st = (cudaStream_t)stream_;
#define LNF(V) launch_pdl(k_ln_fwd_vec<V>, dim3(min(ceil_div(ceil_div(rows, rpw), 8), resident_blocks(k_ln_fwd_vec<V>, 256, 0))), dim3(256), 0, st, (const bf16_t*)x, xm, (bf16_t*)y, ym, (const bf16_t*)residual, gamma, beta, mean, rstd, rows, C, lpr, eps)
    if (vpl == 1) LNF(1); else if (vpl == 2) LNF(2); else if (vpl == 3) LNF(3); else LNF(4);
#undef LNF
    LOTUS_CUDA_OK();
    return 0;
  }
  k_ln_fwd<T><<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream_>>>((const T*)x, xm, (T*)y, ym, (const T*)residual, gamma, beta,
                                                                   mean, rstd, rows, C, eps);
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_layernorm_fwd_f32(const float* x, int x_inner, int64_t x_outer_stride, int64_t x_inner_stride,
                                      float* y, int y_inner, int64_t y_outer_stride, int64_t y_inner_stride,
                                      const float* residual, const float* gamma, const float* beta, float* mean,
                                      float* rstd, int64_t rows, int C, float eps, void* stream_) {
  return ln_fwd_impl<float>(x, x_inner, x_outer_stride, x_inner_stride, y, y_inner, y_outer_stride, y_inner_stride, residual,
                            gamma, beta, mean, rstd, rows, C, eps, stream_);
}
// bf16 activations (x, y, residual), fp32 parameters and statistics
LOTUS_API int lotus_layernorm_fwd_bf16(const void* x, int x_inner, int64_t x_outer_stride, int64_t x_inner_stride, void* y,
                                       int y_inner, int64_t y_outer_stride, int64_t y_inner_stride, const void* residual,
                                       const float* gamma, const float* beta, float* mean, float* rstd, int64_t rows, int C,
                                       float eps, void* stream_) {
  return ln_fwd_impl<bf16_t>(x, x_inner, x_outer_stride, x_inner_stride, y, y_inner, y_outer_stride, y_inner_stride, residual,
                             gamma, beta, mean, rstd, rows, C, eps, stream_);
}

template <typename TDY, typename T>
static int ln_bwd_impl(const void* dy, int dy_inner, int64_t dy_outer_stride, int64_t dy_inner_stride, const void* x,
                       int x_inner, int64_t x_outer_stride, int64_t x_inner_stride, void* dx, int dx_inner,
                       int64_t dx_outer_stride, int64_t dx_inner_stride, const float* gamma, const float* mean,
                       const float* rstd, float* dgamma, float* dbeta, int64_t rows, int C, int accumulate_dx, void* stream_,
                       void* dx_masked = nullptr, float drop_p = 0.f, uint64_t seed = 0, uint32_t drop_sid = 0) {
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  RowMap dym{dy_inner, dy_outer_stride, dy_inner_stride}, xm{x_inner, x_outer_stride, x_inner_stride},
      dxm{dx_inner, dx_outer_stride, dx_inner_stride};
  int blocks = min(ceil_div(rows, 64), kNumSM * 4);
  int rpb = ceil_div(rows, blocks);
  blocks = ceil_div(rows, rpb);
  size_t smem = (size_t)2 * C * 4;
  LOTUS_CHECK_ARG(C <= 32 * 32, "lotus_layernorm_bwd: C=%d > 1024", C);
  if (sizeof(T) == 2 && C % 8 == 0 && vec_ok(dym) && vec_ok(xm) && vec_ok(dxm) && aligned16(dy) && aligned16(x) && aligned16(dx) &&
      aligned16(gamma)) {
    int lpr = pick_lpr(C), vpl = ceil_div(C >> 3, lpr), rpw = 32 / lpr;
#define LNB(V, U) launch_pdl(k_ln_bwd_vec<V, U, TDY>, dim3(min(ceil_div(ceil_div(rows, rpw * U), 8), resident_blocks(k_ln_bwd_vec<V, U, TDY>, 256, smem))), dim3(256), smem, st, (const TDY*)dy, dym, (const bf16_t*)x, xm, (bf16_t*)dx, dxm, gamma, mean, rstd, dgamma, dbeta, rows, C, lpr, accumulate_dx, (bf16_t*)dx_masked, drop_p, seed, drop_sid)
    if (vpl == 1) LNB(1, 2); else if (vpl == 2) LNB(2, 1); else if (vpl == 3) LNB(3, 1); else LNB(4, 1);
#undef LNB
    LOTUS_CUDA_OK();
    return 0;
  }
  LOTUS_CHECK_ARG(dx_masked == nullptr, "layernorm_bwd: the masked second output needs the vectorised path (C %% 8 == 0, 16-byte aligned rows)");
#define LN_BWD(CPL)                                                                                                   \
  k_ln_bwd<CPL, TDY, T><<<blocks, 256, smem, st>>>((const TDY*)dy, dym, (const T*)x, xm, (T*)dx, dxm, gamma, mean, rstd, \
                                                   dgamma, dbeta, rows, C, rpb, accumulate_dx)
  if (C <= 32) LN_BWD(1);
  else if (C <= 64) LN_BWD(2);
  else if (C <= 128) LN_BWD(4);
  else if (C <= 256) LN_BWD(8);
  else if (C <= 512) LN_BWD(16);
  else LN_BWD(32);
#undef LN_BWD
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_layernorm_bwd_f32(const float* dy, int dy_inner, int64_t dy_outer_stride, int64_t dy_inner_stride,
                                      const float* x, int x_inner, int64_t x_outer_stride, int64_t x_inner_stride,
                                      float* dx, int dx_inner, int64_t dx_outer_stride, int64_t dx_inner_stride,
                                      const float* gamma, const float* mean, const float* rstd, float* dgamma,
                                      float* dbeta, int64_t rows, int C, int accumulate_dx, void* stream_) {
  return ln_bwd_impl<float, float>(dy, dy_inner, dy_outer_stride, dy_inner_stride, x, x_inner, x_outer_stride, x_inner_stride, dx,
                                   dx_inner, dx_outer_stride, dx_inner_stride, gamma, mean, rstd, dgamma, dbeta, rows, C,
                                   accumulate_dx, stream_);
}
// x, dx bf16; dy bf16 (dy_fp32 = 0) or fp32 (dy_fp32 = 1: gradients coming out of the attention kernels)
LOTUS_API int lotus_layernorm_bwd_bf16(const void* dy, int dy_fp32, int dy_inner, int64_t dy_outer_stride,
                                       int64_t dy_inner_stride, const void* x, int x_inner, int64_t x_outer_stride,
                                       int64_t x_inner_stride, void* dx, int dx_inner, int64_t dx_outer_stride,
                                       int64_t dx_inner_stride, const float* gamma, const float* mean, const float* rstd,
                                       float* dgamma, float* dbeta, int64_t rows, int C, int accumulate_dx, void* stream_) {
  if (dy_fp32)
    return ln_bwd_impl<float, bf16_t>(dy, dy_inner, dy_outer_stride, dy_inner_stride, x, x_inner, x_outer_stride, x_inner_stride,
                                      dx, dx_inner, dx_outer_stride, dx_inner_stride, gamma, mean, rstd, dgamma, dbeta, rows, C,
                                      accumulate_dx, stream_);
  return ln_bwd_impl<bf16_t, bf16_t>(dy, dy_inner, dy_outer_stride, dy_inner_stride, x, x_inner, x_outer_stride, x_inner_stride,
                                     dx, dx_inner, dx_outer_stride, dx_inner_stride, gamma, mean, rstd, dgamma, dbeta, rows, C,
                                     accumulate_dx, stream_);
}

// contiguous [rows, C] form that also emits dropout_mask(dx) for the next layer's output dropout (drop_p, seed, drop_stream as in
// lotus_dropout_bf16): saves the separate masking pass over the residual-stream gradient.  x, dy, dx, dx_masked bf16.
LOTUS_API int lotus_layernorm_bwd_mask_bf16(const void* dy, const void* x, void* dx, void* dx_masked, const float* gamma, const float* mean,
                                            const float* rstd, float* dgamma, float* dbeta, int64_t rows, int C, int accumulate_dx,
                                            float drop_p, uint64_t seed, uint32_t drop_stream, void* stream_) {
  LOTUS_CHECK_ARG(dx_masked != nullptr && drop_p > 0.f && drop_p < 1.f && aligned16(dx_masked), "lotus_layernorm_bwd_mask_bf16: needs a masked output and 0 < p < 1");
  return ln_bwd_impl<bf16_t, bf16_t>(dy, 1, C, 0, x, 1, C, 0, dx, 1, C, 0, gamma, mean, rstd, dgamma, dbeta, rows, C, accumulate_dx, stream_,
                                     dx_masked, drop_p, seed, drop_stream);
}

// ---------------------------------------------------------------- BatchNorm (+GELU)
// column statistics: grid (ceil(C/32), row_blocks); each block reduces a [rows_chunk x 32] panel
template <typename T>
__global__ void k_col_sum(const T* __restrict__ x, int64_t rows, int C, const float* __restrict__ shift /*C or null*/,
                          int square, float* __restrict__ out /*C, pre-zeroed*/, int rows_per_block) {
  __shared__ float red[8][33];
  int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;  // 32 x 8
  int c = blockIdx.x * 32 + cx;
  int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  if (c < C) {
    float sh = shift ? shift[c] : 0.f;
    for (int64_t r = r0 + ry; r < r1; r += 8) {
      float v = ldf(x, r * C + c) - sh;
      s += square ? v * v : v;
    }
  }
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][cx];
    atomicAdd(&out[c], t);
  }
}

static void col_launch_cfg(int64_t rows, int C, dim3* grid, int* rpb) {
  int cb = ceil_div(C, 32);
  int rb = max(1, min(ceil_div(rows, 64), (kNumSM * 8) / cb + 1));
  *rpb = ceil_div(rows, rb);
  rb = ceil_div(rows, *rpb);
  *grid = dim3(cb, rb);
}


// 8 bf16 per thread (16-byte accesses); same op codes and dropout mask as k_pointwise
__global__ void __launch_bounds__(256) k_pointwise_vec(const bf16_t* __restrict__ x, const bf16_t* __restrict__ a, bf16_t* __restrict__ y,
                                                       int64_t nvec, int op, float alpha, float drop_p, uint64_t seed, uint32_t sid) {
  pdl_wait();                                                    // programmatic dependent launch (common.cuh)
  pdl_launch_dependents();
  const uint32_t thr = drop_thr(drop_p);
  const float ks = drop_p > 0.f ? drop_scale(thr) : 1.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8], av[8], o[8], m[8];
    load8(x + i * 8, v);
    if (op == 1 || op == 3 || op == 4 || op == 8) load8(a + i * 8, av);
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = 1.f;
    if (drop_p > 0.f && op < 4) {
      uint32_t h0 = drop_hash4(seed, sid, (uint64_t)i * 8), h1 = drop_hash4(seed, sid, (uint64_t)i * 8 + 4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        m[e] = (((h0 >> (8 * e)) & 0xFFu) >= thr) ? ks : 0.f;
        m[4 + e] = (((h1 >> (8 * e)) & 0xFFu) >= thr) ? ks : 0.f;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      switch (op) {
        case 0: o[e] = gelu_f(v[e]) * m[e]; break;
        case 1: o[e] = av[e] * m[e] * gelu_grad_f(v[e]); break;
        case 2: o[e] = (v[e] > 0.f ? v[e] : v[e] * alpha) * m[e]; break;
        case 3: o[e] = av[e] * m[e] * (v[e] > 0.f ? 1.f : alpha); break;
        case 4: o[e] = av[e] + v[e]; break;
        case 5: o[e] = v[e] * alpha; break;
        case 7: o[e] = v[e] / (1.f + __expf(-v[e])); break;
        case 8: { float sg = 1.f / (1.f + __expf(-v[e])); o[e] = av[e] * sg * (1.f + v[e] * (1.f - sg)); break; }
        default: o[e] = v[e]; break;
      }
    }
    *(uint4*)(y + i * 8) = pack8(o);
  }
}
__global__ void __launch_bounds__(256) k_dropout_vec(const bf16_t* __restrict__ x, bf16_t* __restrict__ y, int64_t nvec, float p,
                                                     uint64_t seed, uint32_t sid) {
  pdl_wait();                                                    // programmatic dependent launch (common.cuh)
  pdl_launch_dependents();
  const uint32_t thr = drop_thr(p);
  const float ks = drop_scale(thr);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    load8(x + i * 8, v);
    uint32_t h0 = drop_hash4(seed, sid, (uint64_t)i * 8), h1 = drop_hash4(seed, sid, (uint64_t)i * 8 + 4);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[e] = (((h0 >> (8 * e)) & 0xFFu) >= thr) ? v[e] * ks : 0.f;
      v[4 + e] = (((h1 >> (8 * e)) & 0xFFu) >= thr) ? v[4 + e] * ks : 0.f;
    }
    *(uint4*)(y + i * 8) = pack8(v);
  }
}
// column sums of a bf16 [rows, C] matrix, C % 8 == 0: thread = (row group, 8-channel vector); block partials via shared memory
__global__ void __launch_bounds__(256) k_col_sum_vec(const bf16_t* __restrict__ x, int64_t rows, int C, float* __restrict__ out,
                                                     int rows_per_block) {
  extern __shared__ float sred[];          // [C]
  const int nvec = C >> 3;                 // vectors per row
  const int tpr = min(nvec, 256);          // threads per row
  const int rg = 256 / tpr;                // row groups per block
  const int tv = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  for (int c = threadIdx.x; c < C; c += blockDim.x) sred[c] = 0.f;
  __syncthreads();
  int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  if (tr < rg) {
    for (int v0 = tv; v0 < nvec; v0 += tpr) {
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      for (int64_t r = r0 + tr; r < r1; r += rg) {
        float v[8];
        load8(x + r * C + v0 * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += v[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&sred[v0 * 8 + e], acc[e]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(&out[c], sred[c]);
}

template <typename T>
static int colsum_impl(const void* x, int64_t rows, int C, float* out_accum, void* stream_) {
  if (rows <= 0) return 0;
  if (sizeof(T) == 2 && C % 8 == 0 && C <= 8192 && aligned16(x)) {
    int blocks = max(1, min(ceil_div(rows, 64), kNumSM * 4));
    int rpbv = ceil_div(rows, blocks);
    blocks = ceil_div(rows, rpbv);
    k_col_sum_vec<<<blocks, 256, (size_t)C * 4, (cudaStream_t)stream_>>>((const bf16_t*)x, rows, C, out_accum, rpbv);
    LOTUS_CUDA_OK();
    return 0;
  }
  dim3 grid; int rpb;
  col_launch_cfg(rows, C, &grid, &rpb);
  k_col_sum<T><<<grid, 256, 0, (cudaStream_t)stream_>>>((const T*)x, rows, C, nullptr, 0, out_accum, rpb);
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_colsum_f32(const float* x, int64_t rows, int C, float* out_accum, void* stream_) {
  return colsum_impl<float>(x, rows, C, out_accum, stream_);
}
LOTUS_API int lotus_colsum_bf16(const void* x, int64_t rows, int C, float* out_accum, void* stream_) {
  return colsum_impl<bf16_t>(x, rows, C, out_accum, stream_);
}

__global__ void k_bn_finalize(float* __restrict__ sum /*C: in sum, out mean*/, float* __restrict__ sq, int64_t rows,
                              int C, float eps, float momentum, float* __restrict__ rstd, float* __restrict__ run_mean,
                              float* __restrict__ run_var, int phase) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (phase == 0) {
    sum[c] = sum[c] / rows;  // mean
  } else {
    float var = sq[c] / rows;  // biased
    rstd[c] = rsqrtf(var + eps);
    if (run_mean) {
      run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * sum[c];
      float unb = rows > 1 ? var * ((float)rows / (float)(rows - 1)) : var;
      run_var[c] = (1.f - momentum) * run_var[c] + momentum * unb;
    }
  }
}

template <typename TX, typename TY>
__global__ void k_bn_apply(const TX* __restrict__ x, TY* __restrict__ y, const float* __restrict__ mean,
                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                           const float* __restrict__ beta, int64_t total, int C, int act) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float z = (ldf(x, i) - mean[c]) * rstd[c] * gamma[c] + beta[c];
    stf(y, i, act ? gelu_f(z) : z);
  }
}


// ---------------------------------------------------------------- fused BatchNorm [+GELU], bf16 path: ONE cooperative launch
// The multi-kernel form above (memset x2, column sums x2, finalize x2, apply) costs 45 us on [96652, 64] where the three passes over
// the tensor are worth 5 us.  Here a grid of resident blocks (cooperative launch: co-scheduled, so the grid barrier cannot deadlock)
// does   zero the accumulators | barrier | pivot-shifted column sums of x and x^2 | barrier | per-channel scale / shift | apply,
// with 16-byte accesses (8 channels per thread) and the second read of x served by L2.  The variance comes from sums shifted by the
// first row (pivot p[c] = x[0][c]):  var = E[(x - p)^2] - E[x - p]^2, which keeps the cancellation harmless when |mean| >> std.
#include <cooperative_groups.h>
#include "tc_common.cuh"
namespace cg = cooperative_groups;
__device__ __forceinline__ void store8(bf16_t* p, const float* v) { *(uint4*)p = pack8(v); }
__device__ __forceinline__ void store8(float* p, const float* v) {
  *(float4*)p = make_float4(v[0], v[1], v[2], v[3]);
  *(float4*)(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
// GELU of the bf16 path: the same function as the GEMM epilogues (tc_common.cuh gelu_fast / gelu_grad_fast)
__device__ __forceinline__ float gelu_na(float x) { return tc::gelu_fast(x, 0.5f); }
__device__ __forceinline__ float gelu_grad_na(float x) { return tc::gelu_grad_fast(x, 0.5f); }

template <typename TX>
__global__ void __launch_bounds__(256) k_bn_fwd_fused(const TX* __restrict__ x, bf16_t* __restrict__ y, int64_t rows, int C,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float* __restrict__ run_mean, float* __restrict__ run_var,
                                                      float* __restrict__ save_mean, float* __restrict__ save_rstd,
                                                      float* __restrict__ acc /*[2][C]*/, int train, float eps, float momentum, int act) {
  __shared__ float sA[1024], sB[1024];
  cg::grid_group grid = cg::this_grid();
  const int nvec = C >> 3, tpr = nvec, rg = 256 / tpr;
  const int tv = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  const bool active = tr < rg;
  const int64_t stride = (int64_t)gridDim.x * rg;
  const int64_t rfirst = (int64_t)blockIdx.x * rg + tr;
  if (train) {
    for (int c = blockIdx.x * 256 + threadIdx.x; c < 2 * C; c += gridDim.x * 256) acc[c] = 0.f;
    for (int c = threadIdx.x; c < 2 * C; c += 256) { if (c < C) sA[c] = 0.f; else sB[c - C] = 0.f; }
    float sq[2][8], pv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sq[0][e] = 0.f; sq[1][e] = 0.f; pv[e] = 0.f; }
    if (active) load8(x + tv * 8, pv);                           // pivot: row 0
    if (active) {
      int64_t r = rfirst;
      for (; r + 3 * stride < rows; r += 4 * stride) {           // four rows in flight
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) load8(x + (r + u * stride) * C + tv * 8, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int e = 0; e < 8; ++e) { const float d = v[u][e] - pv[e]; sq[0][e] += d; sq[1][e] = fmaf(d, d, sq[1][e]); }
      }
      for (; r < rows; r += stride) {
        float v[8];
        load8(x + r * C + tv * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = v[e] - pv[e]; sq[0][e] += d; sq[1][e] = fmaf(d, d, sq[1][e]); }
      }
    }
    __syncthreads();
    // both partial sets go through one [2][C] view: sA = sums, sB = squares
    {
      const int lane = threadIdx.x & 31;
      const bool pow2 = (tpr & (tpr - 1)) == 0 && tpr < 32;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float a = active ? sq[0][e] : 0.f, b = active ? sq[1][e] : 0.f;
        if (pow2) for (int o = 16; o >= tpr; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (active && (!pow2 || lane < tpr)) { atomicAdd(&sA[tv * 8 + e], a); atomicAdd(&sB[tv * 8 + e], b); }
      }
    }
    grid.sync();                                                 // accumulators zeroed everywhere (and this block's partials complete)
    for (int c = threadIdx.x; c < C; c += 256) { atomicAdd(&acc[c], sA[c]); atomicAdd(&acc[C + c], sB[c]); }
    grid.sync();
  }
  // per-channel scale / shift
  const float invn = 1.f / (float)rows;
  for (int c = threadIdx.x; c < C; c += 256) {
    float mean, rs;
    if (train) {
      const float m = __ldcg(&acc[c]) * invn;
      const float var = fmaxf(__ldcg(&acc[C + c]) * invn - m * m, 0.f);
      mean = ldf(x, c) + m;
      rs = rsqrtf(var + eps);
      if (blockIdx.x == 0 && run_mean) {
        run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
        const float unb = rows > 1 ? var * ((float)rows / (float)(rows - 1)) : var;
        run_var[c] = (1.f - momentum) * run_var[c] + momentum * unb;
      }
    } else {
      mean = run_mean[c];
      rs = rsqrtf(run_var[c] + eps);
    }
    if (blockIdx.x == 0) { save_mean[c] = mean; save_rstd[c] = rs; }
    const float sc = rs * gamma[c];
    sA[c] = sc;
    sB[c] = beta[c] - mean * sc;
  }
  __syncthreads();
  if (!active) return;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { sc[e] = sA[tv * 8 + e]; sh[e] = sB[tv * 8 + e]; }
  int64_t r = rfirst;
  for (; r + 3 * stride < rows; r += 4 * stride) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) load8(x + (r + u * stride) * C + tv * 8, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float z = fmaf(v[u][e], sc[e], sh[e]); v[u][e] = act ? gelu_na(z) : z; }
      store8(y + (r + u * stride) * C + tv * 8, v[u]);
    }
  }
  for (; r < rows; r += stride) {
    float v[8];
    load8(x + r * C + tv * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float z = fmaf(v[e], sc[e], sh[e]); v[e] = act ? gelu_na(z) : z; }
    store8(y + r * C + tv * 8, v);
  }
}

// backward of y = act(bn(x)), training statistics: zero | barrier | sum dz, sum dz * xhat | barrier | dx; the batch sums are left in
// dgamma_batch / dbeta_batch and (optionally) added to the parameter-gradient accumulators by block 0
template <typename TX>
__global__ void __launch_bounds__(256) k_bn_bwd_fused(const bf16_t* __restrict__ dy, const TX* __restrict__ x, int64_t rows, int C,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ mean, const float* __restrict__ rstd, TX* __restrict__ dx,
                                                      float* __restrict__ dgamma_batch, float* __restrict__ dbeta_batch,
                                                      float* __restrict__ dgamma_acc, float* __restrict__ dbeta_acc, int act) {
  __shared__ float sA[1024], sB[1024];
  cg::grid_group grid = cg::this_grid();
  const int nvec = C >> 3, tpr = nvec, rg = 256 / tpr;
  const int tv = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  const bool active = tr < rg;
  const int64_t stride = (int64_t)gridDim.x * rg;
  const int64_t rfirst = (int64_t)blockIdx.x * rg + tr;
  for (int c = blockIdx.x * 256 + threadIdx.x; c < C; c += gridDim.x * 256) { dgamma_batch[c] = 0.f; dbeta_batch[c] = 0.f; }
  for (int c = threadIdx.x; c < C; c += 256) { sA[c] = 0.f; sB[c] = 0.f; }
  float mu[8], rs[8], g[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { mu[e] = 0.f; rs[e] = 0.f; g[e] = 0.f; b[e] = 0.f; }
  if (active) { load8(mean + tv * 8, mu); load8(rstd + tv * 8, rs); load8(gamma + tv * 8, g); load8(beta + tv * 8, b); }
  {
    float sg[8], sb[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sg[e] = 0.f; sb[e] = 0.f; }
    if (active) {
      int64_t r = rfirst;
      for (; r + stride < rows; r += 2 * stride) {               // two rows (four vectors) in flight
        float xv[2][8], dv[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u) { load8(x + (r + u * stride) * C + tv * 8, xv[u]); load8(dy + (r + u * stride) * C + tv * 8, dv[u]); }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float xh = (xv[u][e] - mu[e]) * rs[e];
            float dz = dv[u][e];
            if (act) dz *= gelu_grad_na(fmaf(xh, g[e], b[e]));
            sg[e] = fmaf(dz, xh, sg[e]); sb[e] += dz;
          }
      }
      for (; r < rows; r += stride) {
        float xv[8], dv[8];
        load8(x + r * C + tv * 8, xv); load8(dy + r * C + tv * 8, dv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (xv[e] - mu[e]) * rs[e];
          float dz = dv[e];
          if (act) dz *= gelu_grad_na(fmaf(xh, g[e], b[e]));
          sg[e] = fmaf(dz, xh, sg[e]); sb[e] += dz;
        }
      }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const bool pow2 = (tpr & (tpr - 1)) == 0 && tpr < 32;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float a = active ? sg[e] : 0.f, bb = active ? sb[e] : 0.f;
      if (pow2) for (int o = 16; o >= tpr; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); bb += __shfl_xor_sync(0xffffffffu, bb, o); }
      if (active && (!pow2 || lane < tpr)) { atomicAdd(&sA[tv * 8 + e], a); atomicAdd(&sB[tv * 8 + e], bb); }
    }
  }
  grid.sync();
  for (int c = threadIdx.x; c < C; c += 256) { atomicAdd(&dgamma_batch[c], sA[c]); atomicAdd(&dbeta_batch[c], sB[c]); }
  grid.sync();
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += 256) {
      if (dgamma_acc) dgamma_acc[c] += __ldcg(&dgamma_batch[c]);
      if (dbeta_acc) dbeta_acc[c] += __ldcg(&dbeta_batch[c]);
    }
  }
  if (!active) return;
  const float invn = 1.f / (float)rows;
  float k1[8], k2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { k2[e] = __ldcg(&dgamma_batch[tv * 8 + e]) * invn; k1[e] = __ldcg(&dbeta_batch[tv * 8 + e]) * invn; }
  int64_t r = rfirst;
  for (; r + stride < rows; r += 2 * stride) {
    float xv[2][8], dv[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u) { load8(x + (r + u * stride) * C + tv * 8, xv[u]); load8(dy + (r + u * stride) * C + tv * 8, dv[u]); }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = (xv[u][e] - mu[e]) * rs[e];
        float dz = dv[u][e];
        if (act) dz *= gelu_grad_na(fmaf(xh, g[e], b[e]));
        xv[u][e] = g[e] * rs[e] * (dz - k1[e] - xh * k2[e]);
      }
      store8(dx + (r + u * stride) * C + tv * 8, xv[u]);
    }
  }
  for (; r < rows; r += stride) {
    float xv[8], dv[8];
    load8(x + r * C + tv * 8, xv); load8(dy + r * C + tv * 8, dv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xh = (xv[e] - mu[e]) * rs[e];
      float dz = dv[e];
      if (act) dz *= gelu_grad_na(fmaf(xh, g[e], b[e]));
      xv[e] = g[e] * rs[e] * (dz - k1[e] - xh * k2[e]);
    }
    store8(dx + r * C + tv * 8, xv);
  }
}

static inline bool bn_fused_ok(int64_t rows, int C, const void* a, const void* b, const void* c) {
  return C % 8 == 0 && C <= 1024 && C >= 8 && (C >> 3) <= 256 && rows >= 1 && aligned16(a) && aligned16(b) && (!c || aligned16(c));
}
template <typename TX>
static int bn_fwd_fused_launch(const void* x, void* y, int64_t rows, int C, const float* gamma, const float* beta, float* run_mean,
                               float* run_var, float* save_mean, float* save_rstd, float* acc, int train, float eps, float momentum,
                               int act, cudaStream_t st) {
  const int rg = 256 / (C >> 3);
  int blocks = (int)min((int64_t)resident_blocks(k_bn_fwd_fused<TX>, 256, 0), (rows + rg - 1) / rg);
  const TX* xp = (const TX*)x; bf16_t* yp = (bf16_t*)y;
  void* args[] = {&xp, &yp, &rows, &C, &gamma, &beta, &run_mean, &run_var, &save_mean, &save_rstd, &acc, &train, &eps, &momentum, &act};
  LOTUS_CUDA_CALL(cudaLaunchCooperativeKernel((const void*)k_bn_fwd_fused<TX>, dim3(blocks), dim3(256), args, 0, st));
  return 0;
}
template <typename TX>
static int bn_bwd_fused_launch(const void* dy, const void* x, int64_t rows, int C, const float* gamma, const float* beta,
                               const float* mean, const float* rstd, void* dx, float* dgb, float* dbb, float* dga, float* dba, int act,
                               cudaStream_t st) {
  const int rg = 256 / (C >> 3);
  int blocks = (int)min((int64_t)resident_blocks(k_bn_bwd_fused<TX>, 256, 0), (rows + rg - 1) / rg);
  const bf16_t* dyp = (const bf16_t*)dy; const TX* xp = (const TX*)x; TX* dxp = (TX*)dx;
  void* args[] = {&dyp, &xp, &rows, &C, &gamma, &beta, &mean, &rstd, &dxp, &dgb, &dbb, &dga, &dba, &act};
  LOTUS_CUDA_CALL(cudaLaunchCooperativeKernel((const void*)k_bn_bwd_fused<TX>, dim3(blocks), dim3(256), args, 0, st));
  return 0;
}

// training: batch statistics (two-pass: mean, then centred second moment), running stats updated in place.
// eval (train=0): mean/rstd derived from the running stats.  save_mean/save_rstd (C each) feed the backward.
template <typename TX, typename TY>
static int bn_fwd_impl(const void* x, void* y, int64_t rows, int C, const float* gamma, const float* beta, float* run_mean,
                       float* run_var, float* save_mean, float* save_rstd, float* scratch_C, int train, float eps,
                       float momentum, int act_gelu, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (rows <= 0) return 0;
  if (train) {
    dim3 grid; int rpb;
    col_launch_cfg(rows, C, &grid, &rpb);
    LOTUS_CUDA_CALL(cudaMemsetAsync(save_mean, 0, (size_t)C * 4, st));
    LOTUS_CUDA_CALL(cudaMemsetAsync(scratch_C, 0, (size_t)C * 4, st));
    k_col_sum<TX><<<grid, 256, 0, st>>>((const TX*)x, rows, C, nullptr, 0, save_mean, rpb);
    k_bn_finalize<<<ceil_div(C, 128), 128, 0, st>>>(save_mean, scratch_C, rows, C, eps, momentum, save_rstd, nullptr, nullptr, 0);
    k_col_sum<TX><<<grid, 256, 0, st>>>((const TX*)x, rows, C, save_mean, 1, scratch_C, rpb);
    k_bn_finalize<<<ceil_div(C, 128), 128, 0, st>>>(save_mean, scratch_C, rows, C, eps, momentum, save_rstd, run_mean, run_var, 1);
  } else {
    // mean := running_mean ; rstd := rsqrt(running_var + eps)
    LOTUS_CUDA_CALL(cudaMemcpyAsync(save_mean, run_mean, (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
    LOTUS_CUDA_CALL(cudaMemcpyAsync(scratch_C, run_var, (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
    k_bn_finalize<<<ceil_div(C, 128), 128, 0, st>>>(save_mean, scratch_C, 1, C, eps, 0.f, save_rstd, nullptr, nullptr, 1);
  }
  int64_t total = rows * C;
  k_bn_apply<TX, TY><<<min(ceil_div(total, 256), kNumSM * 16), 256, 0, st>>>((const TX*)x, (TY*)y, save_mean, save_rstd, gamma, beta,
                                                                             total, C, act_gelu);
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_batchnorm_fwd_f32(const float* x, float* y, int64_t rows, int C, const float* gamma,
                                      const float* beta, float* run_mean, float* run_var, float* save_mean,
                                      float* save_rstd, float* scratch_C, int train, float eps, float momentum,
                                      int act_gelu, void* stream_) {
  return bn_fwd_impl<float, float>(x, y, rows, C, gamma, beta, run_mean, run_var, save_mean, save_rstd, scratch_C, train, eps,
                                   momentum, act_gelu, stream_);
}
// y bf16; x bf16 (x_fp32 = 0) or fp32 (x_fp32 = 1: the stem, whose conv runs in fp32)
LOTUS_API int lotus_batchnorm_fwd_bf16(const void* x, int x_fp32, void* y, int64_t rows, int C, const float* gamma,
                                       const float* beta, float* run_mean, float* run_var, float* save_mean,
                                       float* save_rstd, float* scratch_C, int train, float eps, float momentum,
                                       int act_gelu, void* stream_) {
  // scratch_C: 2 * C floats on this path (column sums and sums of squares)
  if (bn_fused_ok(rows, C, x, y, nullptr) && aligned16(gamma) && aligned16(beta)) {
    if (x_fp32) return bn_fwd_fused_launch<float>(x, y, rows, C, gamma, beta, run_mean, run_var, save_mean, save_rstd, scratch_C, train, eps, momentum, act_gelu, (cudaStream_t)stream_);
    return bn_fwd_fused_launch<bf16_t>(x, y, rows, C, gamma, beta, run_mean, run_var, save_mean, save_rstd, scratch_C, train, eps, momentum, act_gelu, (cudaStream_t)stream_);
  }
  if (x_fp32)
    return bn_fwd_impl<float, bf16_t>(x, y, rows, C, gamma, beta, run_mean, run_var, save_mean, save_rstd, scratch_C, train, eps,
                                      momentum, act_gelu, stream_);
  return bn_fwd_impl<bf16_t, bf16_t>(x, y, rows, C, gamma, beta, run_mean, run_var, save_mean, save_rstd, scratch_C, train, eps,
                                     momentum, act_gelu, stream_);
}

// backward of y = act(bn(x)).  pass 1: dbeta = sum dz, dgamma = sum dz*xhat (dz = dy * act'(z)); pass 2: dx.
template <typename TX, typename TY>
__global__ void k_bn_bwd_reduce(const TY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int64_t rows, int C, int act,
                                float* __restrict__ dgamma, float* __restrict__ dbeta, int rows_per_block) {
  __shared__ float r1s[8][33], r2s[8][33];
  int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  int c = blockIdx.x * 32 + cx;
  int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float sg = 0.f, sb = 0.f;
  if (c < C) {
    float mu = mean[c], rs = rstd[c], g = gamma[c], b = beta[c];
    for (int64_t r = r0 + ry; r < r1; r += 8) {
      float xh = (ldf(x, r * C + c) - mu) * rs;
      float dz = ldf(dy, r * C + c);
      if (act) dz *= gelu_grad_f(xh * g + b);
      sg += dz * xh; sb += dz;
    }
  }
  r1s[ry][cx] = sg; r2s[ry][cx] = sb;
  __syncthreads();
  if (ry == 0 && c < C) {
    float a = 0.f, b2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a += r1s[i][cx]; b2 += r2s[i][cx]; }
    atomicAdd(&dgamma[c], a); atomicAdd(&dbeta[c], b2);
  }
}
template <typename TX, typename TY>
__global__ void k_bn_bwd_apply(const TY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ mean,
                               const float* __restrict__ rstd, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ dgamma_batch,
                               const float* __restrict__ dbeta_batch, int64_t rows, int C, int act, int train,
                               TX* __restrict__ dx) {
  int64_t total = rows * C;
  float invn = 1.f / (float)rows;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float rs = rstd[c], g = gamma[c];
    float xh = (ldf(x, i) - mean[c]) * rs;
    float dz = ldf(dy, i);
    if (act) dz *= gelu_grad_f(xh * g + beta[c]);
    stf(dx, i, train ? g * rs * (dz - dbeta_batch[c] * invn - xh * dgamma_batch[c] * invn) : g * rs * dz);
  }
}

// dgamma_batch/dbeta_batch: zero-initialised C-float scratch receiving THIS batch's sums (the caller adds
// them to the parameter gradients).
template <typename TX, typename TY>
static int bn_bwd_impl(const void* dy, const void* x, int64_t rows, int C, const float* gamma, const float* beta,
                       const float* save_mean, const float* save_rstd, void* dx, float* dgamma_batch, float* dbeta_batch,
                       int train, int act_gelu, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (rows <= 0) return 0;
  dim3 grid; int rpb;
  col_launch_cfg(rows, C, &grid, &rpb);
  k_bn_bwd_reduce<TX, TY><<<grid, 256, 0, st>>>((const TY*)dy, (const TX*)x, save_mean, save_rstd, gamma, beta, rows, C, act_gelu,
                                                dgamma_batch, dbeta_batch, rpb);
  int64_t total = rows * C;
  k_bn_bwd_apply<TX, TY><<<min(ceil_div(total, 256), kNumSM * 16), 256, 0, st>>>(
      (const TY*)dy, (const TX*)x, save_mean, save_rstd, gamma, beta, dgamma_batch, dbeta_batch, rows, C, act_gelu, train, (TX*)dx);
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_batchnorm_bwd_f32(const float* dy, const float* x, int64_t rows, int C, const float* gamma,
                                      const float* beta, const float* save_mean, const float* save_rstd, float* dx,
                                      float* dgamma_batch, float* dbeta_batch, int train, int act_gelu, void* stream_) {
  return bn_bwd_impl<float, float>(dy, x, rows, C, gamma, beta, save_mean, save_rstd, dx, dgamma_batch, dbeta_batch, train,
                                   act_gelu, stream_);
}
// dy bf16; x and dx bf16 (x_fp32 = 0) or fp32 (x_fp32 = 1)
LOTUS_API int lotus_batchnorm_bwd_bf16(const void* dy, const void* x, int x_fp32, int64_t rows, int C, const float* gamma,
                                       const float* beta, const float* save_mean, const float* save_rstd, void* dx,
                                       float* dgamma_batch, float* dbeta_batch, int train, int act_gelu, void* stream_) {
  if (train && bn_fused_ok(rows, C, x, dy, dx) && aligned16(gamma) && aligned16(beta) && aligned16(save_mean) && aligned16(save_rstd)) {
    if (x_fp32) return bn_bwd_fused_launch<float>(dy, x, rows, C, gamma, beta, save_mean, save_rstd, dx, dgamma_batch, dbeta_batch, nullptr, nullptr, act_gelu, (cudaStream_t)stream_);
    return bn_bwd_fused_launch<bf16_t>(dy, x, rows, C, gamma, beta, save_mean, save_rstd, dx, dgamma_batch, dbeta_batch, nullptr, nullptr, act_gelu, (cudaStream_t)stream_);
  }
  if (x_fp32)
    return bn_bwd_impl<float, bf16_t>(dy, x, rows, C, gamma, beta, save_mean, save_rstd, dx, dgamma_batch, dbeta_batch, train,
                                      act_gelu, stream_);
  return bn_bwd_impl<bf16_t, bf16_t>(dy, x, rows, C, gamma, beta, save_mean, save_rstd, dx, dgamma_batch, dbeta_batch, train,
                                     act_gelu, stream_);
}

// training-mode backward in one launch that also adds the batch sums into the parameter-gradient accumulators (either may be null);
// dgamma_batch / dbeta_batch: C-float scratch each (zeroed by the kernel).  Fails on shapes the fused kernel does not take.
LOTUS_API int lotus_batchnorm_bwd_acc_bf16(const void* dy, const void* x, int x_fp32, int64_t rows, int C, const float* gamma,
                                           const float* beta, const float* save_mean, const float* save_rstd, void* dx,
                                           float* dgamma_batch, float* dbeta_batch, float* dgamma_acc, float* dbeta_acc, int act_gelu,
                                           void* stream_) {
  if (rows <= 0) return 0;
  LOTUS_CHECK_ARG(bn_fused_ok(rows, C, x, dy, dx) && aligned16(gamma) && aligned16(beta) && aligned16(save_mean) && aligned16(save_rstd),
                  "lotus_batchnorm_bwd_acc_bf16: C=%d must be a multiple of 8, <= 1024, 16-byte aligned operands", C);
  if (x_fp32) return bn_bwd_fused_launch<float>(dy, x, rows, C, gamma, beta, save_mean, save_rstd, dx, dgamma_batch, dbeta_batch, dgamma_acc, dbeta_acc, act_gelu, (cudaStream_t)stream_);
  return bn_bwd_fused_launch<bf16_t>(dy, x, rows, C, gamma, beta, save_mean, save_rstd, dx, dgamma_batch, dbeta_batch, dgamma_acc, dbeta_acc, act_gelu, (cudaStream_t)stream_);
}

// ---------------------------------------------------------------- pointwise
// op 0: y = drop(gelu(x))            1: y = drop(a) * gelu'(x)     (x = pre-activation, a = dy)
// op 2: y = drop(leaky(x, alpha))    3: y = drop(a) * leaky'(x)
// op 4: y = a + x                    5: y = x * alpha              6: y = x (type conversion)
// op 7: y = silu(x)                  8: y = a * silu'(x)
// drop(.) is the identity when drop_p == 0; mask index = flat element index (same in fwd and bwd)
template <typename TX, typename TY>
__global__ void k_pointwise(const TX* __restrict__ x, const TX* __restrict__ a, TY* __restrict__ y, int64_t n,
                            int op, float alpha, float drop_p, uint64_t seed, uint32_t sid) {
  const uint32_t thr = drop_thr(drop_p);
  float ks = drop_p > 0.f ? drop_scale(thr) : 1.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = ldf(x, i), o;
    float m = 1.f;
    if (drop_p > 0.f && op < 4) m = drop_keep_e(seed, sid, (uint64_t)i, thr) ? ks : 0.f;
    switch (op) {
      case 0: o = gelu_f(v) * m; break;
      case 1: o = ldf(a, i) * m * gelu_grad_f(v); break;
      case 2: o = (v > 0.f ? v : v * alpha) * m; break;
      case 3: o = ldf(a, i) * m * (v > 0.f ? 1.f : alpha); break;
      case 4: o = ldf(a, i) + v; break;
      case 5: o = v * alpha; break;
      case 7: o = v / (1.f + expf(-v)); break;
      case 8: { float sg = 1.f / (1.f + expf(-v)); o = ldf(a, i) * sg * (1.f + v * (1.f - sg)); break; }
      default: o = v; break;
    }
    stf(y, i, o);
  }
}
LOTUS_API int lotus_pointwise_f32(const float* x, const float* a, float* y, int64_t n, int op, float alpha, float drop_p,
                                  uint64_t seed, uint32_t drop_stream, void* stream_) {
  if (n <= 0) return 0;
  k_pointwise<float, float><<<min(ceil_div(n, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>(x, a, y, n, op, alpha, drop_p,
                                                                                                 seed, drop_stream);
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_pointwise_bf16(const void* x, const void* a, void* y, int64_t n, int op, float alpha, float drop_p,
                                   uint64_t seed, uint32_t drop_stream, void* stream_) {
  if (n <= 0) return 0;
  if (n % 8 == 0 && aligned16(x) && aligned16(y) && (!a || aligned16(a))) {
    launch_pdl(k_pointwise_vec, dim3(min(ceil_div(n / 8, 256), kNumSM * 16)), dim3(256), 0, (cudaStream_t)stream_, 
        (const bf16_t*)x, (const bf16_t*)a, (bf16_t*)y, n / 8, op, alpha, drop_p, seed, drop_stream);
    LOTUS_CUDA_OK();
    return 0;
  }
  k_pointwise<bf16_t, bf16_t><<<min(ceil_div(n, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>(
      (const bf16_t*)x, (const bf16_t*)a, (bf16_t*)y, n, op, alpha, drop_p, seed, drop_stream);
  LOTUS_CUDA_OK();
  return 0;
}
// dtype conversion: dir 0: fp32 -> bf16, 1: bf16 -> fp32
LOTUS_API int lotus_convert(const void* x, void* y, int64_t n, int dir, void* stream_) {
  if (n <= 0) return 0;
  int blocks = min(ceil_div(n, 256), kNumSM * 16);
  if (dir == 0)
    k_pointwise<float, bf16_t><<<blocks, 256, 0, (cudaStream_t)stream_>>>((const float*)x, nullptr, (bf16_t*)y, n, 6, 0.f, 0.f, 0, 0);
  else
    k_pointwise<bf16_t, float><<<blocks, 256, 0, (cudaStream_t)stream_>>>((const bf16_t*)x, nullptr, (float*)y, n, 6, 0.f, 0.f, 0, 0);
  LOTUS_CUDA_OK();
  return 0;
}

// strided fp32 -> bf16 row copy: y[r, :C] (row stride ldy) = x[r, :C] (row stride ldx)
__global__ void k_convert_rows(const float* __restrict__ x, int64_t ldx, bf16_t* __restrict__ y, int64_t ldy, int64_t rows, int C) {
  int64_t total = rows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / C;
    int c = (int)(i - r * C);
    y[r * ldy + c] = __float2bfloat16(x[r * ldx + c]);
  }
}
LOTUS_API int lotus_convert_rows_f32_to_bf16(const float* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int C, void* stream_) {
  if (rows <= 0) return 0;
  k_convert_rows<<<min(ceil_div(rows * C, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>(x, ldx, (bf16_t*)y, ldy, rows, C);
  LOTUS_CUDA_OK();
  return 0;
}

// weight staging for the tensor-core path: w fp32 [A][D][B] -> w_bf16 (same layout) and wt_bf16 [B][D][A]
// (Linear: D = 1 -> plain transpose; SubMConv3d: [Cout][K3][Cin] -> [Cin][K3][Cout])
__global__ void k_cast_transpose(const float* __restrict__ w, bf16_t* __restrict__ wb, bf16_t* __restrict__ wt, int A, int D, int B) {
  __shared__ float tile[32][33];
  int d = blockIdx.z;
  int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    int a = a0 + i, b = b0 + tx;
    float v = 0.f;
    if (a < A && b < B) {
      v = w[((int64_t)a * D + d) * B + b];
      if (wb) wb[((int64_t)a * D + d) * B + b] = __float2bfloat16(v);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  if (wt) {
    for (int i = ty; i < 32; i += 8) {
      int b = b0 + i, a = a0 + tx;
      if (a < A && b < B) wt[((int64_t)b * D + d) * A + a] = __float2bfloat16(tile[tx][i]);
    }
  }
}
LOTUS_API int lotus_cast_transpose_bf16(const float* w, void* w_bf16, void* wt_bf16, int A, int D, int B, void* stream_) {
  if (A <= 0 || B <= 0 || D <= 0) return 0;
  dim3 grid(ceil_div(B, 32), ceil_div(A, 32), D);
  k_cast_transpose<<<grid, 256, 0, (cudaStream_t)stream_>>>(w, (bf16_t*)w_bf16, (bf16_t*)wt_bf16, A, D, B);
  LOTUS_CUDA_OK();
  return 0;
}

// batched form: every weight of the model in ONE launch (the per-weight launches were latency-bound: 114 per step)
struct CastBatch {
  static constexpr int MAX = 128;
  const float* w[MAX];
  bf16_t* wb[MAX];
  bf16_t* wt[MAX];
  int A[MAX], D[MAX], B[MAX];
  int tile_start[MAX + 1];   // prefix sum of 32x32 tiles (x D) per entry
  int n;
};
// 64 x 64 tiles: fp32 rows read as float4 (256-byte segments), both bf16 outputs written as 8-byte vectors (128-byte segments);
// weights whose A or B is not a multiple of 4 (or whose pointers are not vector aligned) take the scalar path of the same kernel
__global__ void __launch_bounds__(256) k_cast_transpose_batch(const __grid_constant__ CastBatch cb) {
  __shared__ float tile[64][65];
  int tid = blockIdx.x;
  int lo = 0, hi = cb.n - 1;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cb.tile_start[mid + 1] > tid) hi = mid; else lo = mid + 1;
  }
  const int e = lo, A = cb.A[e], D = cb.D[e], B = cb.B[e];
  const float* __restrict__ w = cb.w[e];
  bf16_t* __restrict__ wb = cb.wb[e];
  bf16_t* __restrict__ wt = cb.wt[e];
  int local = tid - cb.tile_start[e];
  const int tb = (B + 63) >> 6, ta = (A + 63) >> 6;
  const int d = local / (ta * tb);
  local -= d * ta * tb;
  const int a0 = (local / tb) * 64, b0 = (local % tb) * 64;
  const bool vec = (A % 4 == 0) && (B % 4 == 0) && (((uintptr_t)w & 15) == 0) && (!wb || ((uintptr_t)wb & 7) == 0) && (!wt || ((uintptr_t)wt & 7) == 0);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 16 x 16
  if (vec) {
#pragma unroll
    for (int i = ty; i < 64; i += 16) {
      const int a = a0 + i, b = b0 + 4 * tx;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a < A && b < B) {
        const int64_t o = ((int64_t)a * D + d) * B + b;
        v = *(const float4*)(w + o);
        if (wb) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
          *(uint2*)(wb + o) = make_uint2(*(uint32_t*)&h0, *(uint32_t*)&h1);
        }
      }
      tile[i][4 * tx] = v.x; tile[i][4 * tx + 1] = v.y; tile[i][4 * tx + 2] = v.z; tile[i][4 * tx + 3] = v.w;
    }
    __syncthreads();
    if (wt) {
#pragma unroll
      for (int i = ty; i < 64; i += 16) {
        const int b = b0 + i, a = a0 + 4 * tx;
        if (a < A && b < B) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(tile[4 * tx][i], tile[4 * tx + 1][i]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(tile[4 * tx + 2][i], tile[4 * tx + 3][i]);
          *(uint2*)(wt + ((int64_t)b * D + d) * A + a) = make_uint2(*(uint32_t*)&h0, *(uint32_t*)&h1);
        }
      }
    }
    return;
  }
  const int sx = threadIdx.x & 63, sy = threadIdx.x >> 6;       // 64 x 4
  for (int i = sy; i < 64; i += 4) {
    int a = a0 + i, b = b0 + sx;
    float v = 0.f;
    if (a < A && b < B) {
      v = w[((int64_t)a * D + d) * B + b];
      if (wb) wb[((int64_t)a * D + d) * B + b] = __float2bfloat16(v);
    }
    tile[i][sx] = v;
  }
  __syncthreads();
  if (wt) {
    for (int i = sy; i < 64; i += 4) {
      int b = b0 + i, a = a0 + sx;
      if (a < A && b < B) wt[((int64_t)b * D + d) * A + a] = __float2bfloat16(tile[sx][i]);
    }
  }
}
// host arrays of n entries (w fp32 [A][D][B] -> w_bf16 same layout, wt_bf16 [B][D][A]; either output pointer may be null)
LOTUS_API int lotus_cast_transpose_batch_bf16(const void* const* w, void* const* w_bf16, void* const* wt_bf16, const int* A,
                                              const int* D, const int* B, int n, void* stream_) {
  for (int base = 0; base < n; base += CastBatch::MAX) {
    CastBatch cb;
    cb.n = min(CastBatch::MAX, n - base);
    cb.tile_start[0] = 0;
    for (int i = 0; i < cb.n; ++i) {
      int j = base + i;
      LOTUS_CHECK_ARG(A[j] > 0 && B[j] > 0 && D[j] > 0, "lotus_cast_transpose_batch_bf16: empty weight %d", j);
      cb.w[i] = (const float*)w[j]; cb.wb[i] = (bf16_t*)w_bf16[j]; cb.wt[i] = (bf16_t*)wt_bf16[j];
      cb.A[i] = A[j]; cb.D[i] = D[j]; cb.B[i] = B[j];
      cb.tile_start[i + 1] = cb.tile_start[i] + ceil_div(A[j], 64) * ceil_div(B[j], 64) * D[j];
    }
    if (cb.tile_start[cb.n] > 0) k_cast_transpose_batch<<<cb.tile_start[cb.n], 256, 0, (cudaStream_t)stream_>>>(cb);
  }
  LOTUS_CUDA_OK();
  return 0;
}

// dropout: y = keep(seed, stream_id, i) ? x / (1-p) : 0 — the same call with dy gives the backward
template <typename T>
__global__ void k_dropout(const T* __restrict__ x, T* __restrict__ y, int64_t n, float p, uint64_t seed, uint32_t stream_id) {
  const uint32_t thr = drop_thr(p);
  float sc = drop_scale(thr);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    stf(y, i, drop_keep_e(seed, stream_id, (uint64_t)i, thr) ? ldf(x, i) * sc : 0.f);
}
LOTUS_API int lotus_dropout_f32(const float* x, float* y, int64_t n, float p, uint64_t seed, uint32_t stream_id,
                                void* stream_) {
  if (n <= 0) return 0;
  LOTUS_CHECK_ARG(p >= 0.f && p < 1.f, "lotus_dropout_f32: p=%f", p);
  k_dropout<float><<<min(ceil_div(n, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>(x, y, n, p, seed, stream_id);
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_dropout_bf16(const void* x, void* y, int64_t n, float p, uint64_t seed, uint32_t stream_id, void* stream_) {
  if (n <= 0) return 0;
  LOTUS_CHECK_ARG(p >= 0.f && p < 1.f, "lotus_dropout_bf16: p=%f", p);
  if (n % 8 == 0 && aligned16(x) && aligned16(y)) {
    launch_pdl(k_dropout_vec, dim3(min(ceil_div(n / 8, 256), kNumSM * 16)), dim3(256), 0, (cudaStream_t)stream_, (const bf16_t*)x, (bf16_t*)y, n / 8, p, seed,
                                                                                          stream_id);
    LOTUS_CUDA_OK();
    return 0;
  }
  k_dropout<bf16_t><<<min(ceil_div(n, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>((const bf16_t*)x, (bf16_t*)y, n, p, seed,
                                                                                          stream_id);
  LOTUS_CUDA_OK();
  return 0;
}


// ---------------------------------------------------------------- PDNorm modulation (model.py:293-303)
// y[i, c] = x[i, c] * (1 + ss[b(i), C + c]) + ss[b(i), c]  with ss = Linear(SiLU(context)) per cloud: [B, 2C] = [shift | scale];
// b(i) from the cloud offsets (points of a cloud are contiguous at every stage).
__device__ __forceinline__ int cloud_of_row(const int32_t* __restrict__ off, int B, int64_t r) {
  int lo = 0, hi = B - 1;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (off[mid + 1] > r) hi = mid; else lo = mid + 1;
  }
  return lo;
}
__global__ void k_modulate_fwd(const float* __restrict__ x, const float* __restrict__ ss, const int32_t* __restrict__ off, int B,
                               int64_t rows, int C, float* __restrict__ y) {
  int64_t total = rows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / C;
    int c = (int)(i - r * C);
    const float* s = ss + (int64_t)cloud_of_row(off, B, r) * 2 * C;
    y[i] = x[i] * (1.f + s[C + c]) + s[c];
  }
}
// dx = dy * (1 + scale);  dss[b, c] = sum_i dy;  dss[b, C + c] = sum_i dy * x     (one block column per cloud)
__global__ void k_modulate_bwd(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ ss,
                               const int32_t* __restrict__ off, int C, float* __restrict__ dx, float* __restrict__ dss) {
  __shared__ float s_a[8][33], s_b[8][33];
  int b = blockIdx.x;
  int c = blockIdx.y * 32 + threadIdx.x;
  int r0 = off[b], r1 = off[b + 1];
  float a0 = 0.f, a1 = 0.f;
  if (c < C) {
    float sc = 1.f + ss[(int64_t)b * 2 * C + C + c];
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      float g = dy[(int64_t)r * C + c];
      a0 += g;
      a1 = fmaf(g, x[(int64_t)r * C + c], a1);
      dx[(int64_t)r * C + c] = g * sc;
    }
  }
  s_a[threadIdx.y][threadIdx.x] = a0;
  s_b[threadIdx.y][threadIdx.x] = a1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { a0 += s_a[k][threadIdx.x]; a1 += s_b[k][threadIdx.x]; }
    dss[(int64_t)b * 2 * C + c] = a0;
    dss[(int64_t)b * 2 * C + C + c] = a1;
  }
}
LOTUS_API int lotus_pdnorm_modulate_fwd_f32(const float* x, const float* shift_scale, const int32_t* cloud_off, int B, int64_t rows,
                                            int C, float* y, void* stream_) {
  if (rows <= 0) return 0;
  k_modulate_fwd<<<min(ceil_div(rows * C, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>(x, shift_scale, cloud_off, B, rows, C, y);
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_pdnorm_modulate_bwd_f32(const float* dy, const float* x, const float* shift_scale, const int32_t* cloud_off,
                                            int B, int C, float* dx, float* d_shift_scale, void* stream_) {
  if (B <= 0) return 0;
  k_modulate_bwd<<<dim3(B, ceil_div(C, 32)), dim3(32, 8), 0, (cudaStream_t)stream_>>>(dy, x, shift_scale, cloud_off, C, dx,
                                                                                       d_shift_scale);
  LOTUS_CUDA_OK();
  return 0;
}
