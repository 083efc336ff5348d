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

// ---------------------------------------------------------------- LayerNorm
// y = [residual +] (x - mean) * rstd * gamma + beta ; one warp per row
__global__ void k_ln_fwd(const float* __restrict__ x, RowMap xm, float* __restrict__ y, RowMap ym,
                         const float* __restrict__ residual, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float* __restrict__ mean, float* __restrict__ rstd,
                         int64_t rows, int C, float eps) {
  int lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* xr = x + xm(r);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  float mu = warp_sum(s) / C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { float d = xr[c] - mu; v += d * d; }
  float rs = rsqrtf(warp_sum(v) / C + eps);
  if (lane == 0) { if (mean) mean[r] = mu; if (rstd) rstd[r] = rs; }
  float* yr = y + ym(r);
  const float* rr = residual ? residual + ym(r) : nullptr;
  for (int c = lane; c < C; c += 32) {
    float o = (xr[c] - mu) * rs * gamma[c] + beta[c];
    yr[c] = rr ? rr[c] + o : o;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma ; dgamma += dy*xhat ; dbeta += dy
// `accumulate_dx`: dx += (used when the normalised tensor also feeds a residual branch)
template <int CPL>  // channels per lane upper bound (C <= 32*CPL)
__global__ void k_ln_bwd(const float* __restrict__ dy, RowMap dym, const float* __restrict__ x, RowMap xm,
                         float* __restrict__ dx, RowMap dxm, const float* __restrict__ gamma,
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
    const float* dyr = dy + dym(r);
    const float* xr = x + xm(r);
    float mu = mean[r], rs = rstd[r];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      int c = lane + 32 * i;
      if (c < C) {
        float xh = (xr[c] - mu) * rs, d = dyr[c], g = d * gamma[c];
        s1 += g; s2 += g * xh;
        ag[i] += d * xh; ab[i] += d;
      }
    }
    s1 = warp_sum(s1) / C; s2 = warp_sum(s2) / C;
    float* dxr = dx + dxm(r);
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      int c = lane + 32 * i;
      if (c < C) {
        float xh = (xr[c] - mu) * rs;
        float o = rs * (dyr[c] * gamma[c] - s1 - xh * s2);
        dxr[c] = accumulate_dx ? dxr[c] + o : o;
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

LOTUS_API int lotus_layernorm_fwd_f32(const float* x, int x_inner, int64_t x_outer_stride, int64_t x_inner_stride,
                                      float* y, int y_inner, int64_t y_outer_stride, int64_t y_inner_stride,
                                      const float* residual, const float* gamma, const float* beta, float* mean,
                                      float* rstd, int64_t rows, int C, float eps, void* stream_) {
  if (rows <= 0) return 0;
  RowMap xm{x_inner, x_outer_stride, x_inner_stride}, ym{y_inner, y_outer_stride, y_inner_stride};
  k_ln_fwd<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream_>>>(x, xm, y, ym, residual, gamma, beta, mean, rstd, rows, C, eps);
  LOTUS_CUDA_OK();
  return 0;
}

LOTUS_API int lotus_layernorm_bwd_f32(const float* dy, int dy_inner, int64_t dy_outer_stride, int64_t dy_inner_stride,
                                      const float* x, int x_inner, int64_t x_outer_stride, int64_t x_inner_stride,
                                      float* dx, int dx_inner, int64_t dx_outer_stride, int64_t dx_inner_stride,
                                      const float* gamma, const float* mean, const float* rstd, float* dgamma,
                                      float* dbeta, int64_t rows, int C, int accumulate_dx, void* stream_) {
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  RowMap dym{dy_inner, dy_outer_stride, dy_inner_stride}, xm{x_inner, x_outer_stride, x_inner_stride},
      dxm{dx_inner, dx_outer_stride, dx_inner_stride};
  int blocks = min(ceil_div(rows, 64), kNumSM * 4);
  int rpb = ceil_div(rows, blocks);
  blocks = ceil_div(rows, rpb);
  size_t smem = (size_t)2 * C * 4;
  LOTUS_CHECK_ARG(C <= 32 * 32, "lotus_layernorm_bwd_f32: C=%d > 1024", C);
#define LN_BWD(CPL)                                                                                          \
  k_ln_bwd<CPL><<<blocks, 256, smem, st>>>(dy, dym, x, xm, dx, dxm, gamma, mean, rstd, dgamma, dbeta, rows, C, rpb, \
                                           accumulate_dx)
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

// ---------------------------------------------------------------- BatchNorm (+GELU)
// column statistics: grid (ceil(C/32), row_blocks); each block reduces a [rows_chunk x 32] panel
__global__ void k_col_sum(const float* __restrict__ x, int64_t rows, int C, const float* __restrict__ shift /*C or null*/,
                          int square, float* __restrict__ out /*C, pre-zeroed*/, int rows_per_block) {
  __shared__ float red[8][33];
  int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;  // 32 x 8
  int c = blockIdx.x * 32 + cx;
  int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  if (c < C) {
    float sh = shift ? shift[c] : 0.f;
    for (int64_t r = r0 + ry; r < r1; r += 8) {
      float v = x[r * C + c] - sh;
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

LOTUS_API int lotus_colsum_f32(const float* x, int64_t rows, int C, float* out_accum, void* stream_) {
  if (rows <= 0) return 0;
  dim3 grid; int rpb;
  col_launch_cfg(rows, C, &grid, &rpb);
  k_col_sum<<<grid, 256, 0, (cudaStream_t)stream_>>>(x, rows, C, nullptr, 0, out_accum, rpb);
  LOTUS_CUDA_OK();
  return 0;
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

__global__ void k_bn_apply(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ mean,
                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                           const float* __restrict__ beta, int64_t total, int C, int act) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float z = (x[i] - mean[c]) * rstd[c] * gamma[c] + beta[c];
    y[i] = act ? gelu_f(z) : z;
  }
}

// training: batch statistics (two-pass: mean, then centred second moment), running stats updated in place.
// eval (train=0): mean/rstd derived from the running stats.  save_mean/save_rstd (C each) feed the backward.
LOTUS_API int lotus_batchnorm_fwd_f32(const float* x, float* y, int64_t rows, int C, const float* gamma,
                                      const float* beta, float* run_mean, float* run_var, float* save_mean,
                                      float* save_rstd, float* scratch_C, int train, float eps, float momentum,
                                      int act_gelu, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (rows <= 0) return 0;
  if (train) {
    dim3 grid; int rpb;
    col_launch_cfg(rows, C, &grid, &rpb);
    LOTUS_CUDA_CALL(cudaMemsetAsync(save_mean, 0, (size_t)C * 4, st));
    LOTUS_CUDA_CALL(cudaMemsetAsync(scratch_C, 0, (size_t)C * 4, st));
    k_col_sum<<<grid, 256, 0, st>>>(x, rows, C, nullptr, 0, save_mean, rpb);
    k_bn_finalize<<<ceil_div(C, 128), 128, 0, st>>>(save_mean, scratch_C, rows, C, eps, momentum, save_rstd, nullptr, nullptr, 0);
    k_col_sum<<<grid, 256, 0, st>>>(x, rows, C, save_mean, 1, scratch_C, rpb);
    k_bn_finalize<<<ceil_div(C, 128), 128, 0, st>>>(save_mean, scratch_C, rows, C, eps, momentum, save_rstd, run_mean, run_var, 1);
  } else {
    // mean := running_mean ; rstd := rsqrt(running_var + eps)
    LOTUS_CUDA_CALL(cudaMemcpyAsync(save_mean, run_mean, (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
    LOTUS_CUDA_CALL(cudaMemcpyAsync(scratch_C, run_var, (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
    k_bn_finalize<<<ceil_div(C, 128), 128, 0, st>>>(save_mean, scratch_C, 1, C, eps, 0.f, save_rstd, nullptr, nullptr, 1);
  }
  int64_t total = rows * C;
  k_bn_apply<<<min(ceil_div(total, 256), kNumSM * 16), 256, 0, st>>>(x, y, save_mean, save_rstd, gamma, beta, total, C, act_gelu);
  LOTUS_CUDA_OK();
  return 0;
}

// backward of y = act(bn(x)).  pass 1: dbeta = sum dz, dgamma = sum dz*xhat (dz = dy * act'(z)); pass 2: dx.
__global__ void k_bn_bwd_reduce(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
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
      float xh = (x[r * C + c] - mu) * rs;
      float dz = dy[r * C + c];
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
__global__ void k_bn_bwd_apply(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                               const float* __restrict__ rstd, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ dgamma_batch,
                               const float* __restrict__ dbeta_batch, int64_t rows, int C, int act, int train,
                               float* __restrict__ dx) {
  int64_t total = rows * C;
  float invn = 1.f / (float)rows;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float rs = rstd[c], g = gamma[c];
    float xh = (x[i] - mean[c]) * rs;
    float dz = dy[i];
    if (act) dz *= gelu_grad_f(xh * g + beta[c]);
    dx[i] = train ? g * rs * (dz - dbeta_batch[c] * invn - xh * dgamma_batch[c] * invn) : g * rs * dz;
  }
}

// dgamma_batch/dbeta_batch: zero-initialised C-float scratch receiving THIS batch's sums (the caller adds
// them to the parameter gradients).
LOTUS_API int lotus_batchnorm_bwd_f32(const float* dy, const float* x, int64_t rows, int C, const float* gamma,
                                      const float* beta, const float* save_mean, const float* save_rstd, float* dx,
                                      float* dgamma_batch, float* dbeta_batch, int train, int act_gelu, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (rows <= 0) return 0;
  dim3 grid; int rpb;
  col_launch_cfg(rows, C, &grid, &rpb);
  k_bn_bwd_reduce<<<grid, 256, 0, st>>>(dy, x, save_mean, save_rstd, gamma, beta, rows, C, act_gelu, dgamma_batch, dbeta_batch, rpb);
  int64_t total = rows * C;
  k_bn_bwd_apply<<<min(ceil_div(total, 256), kNumSM * 16), 256, 0, st>>>(dy, x, save_mean, save_rstd, gamma, beta, dgamma_batch,
                                                                       dbeta_batch, rows, C, act_gelu, train, dx);
  LOTUS_CUDA_OK();
  return 0;
}

// ---------------------------------------------------------------- pointwise
// op 0: y = drop(gelu(x))            1: y = drop(a) * gelu'(x)     (x = pre-activation, a = dy)
// op 2: y = drop(leaky(x, alpha))    3: y = drop(a) * leaky'(x)
// op 4: y = a + x                    5: y = x * alpha
// drop(.) is the identity when drop_p == 0; mask index = flat element index (same in fwd and bwd)
__global__ void k_pointwise(const float* __restrict__ x, const float* __restrict__ a, float* __restrict__ y, int64_t n,
                            int op, float alpha, float drop_p, uint64_t seed, uint32_t sid) {
  float ks = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i], o;
    float m = 1.f;
    if (drop_p > 0.f && op < 4) m = dropout_keep(seed, sid ^ (uint32_t)(i >> 32), (uint32_t)i, drop_p) ? ks : 0.f;
    switch (op) {
      case 0: o = gelu_f(v) * m; break;
      case 1: o = a[i] * m * gelu_grad_f(v); break;
      case 2: o = (v > 0.f ? v : v * alpha) * m; break;
      case 3: o = a[i] * m * (v > 0.f ? 1.f : alpha); break;
      case 4: o = a[i] + v; break;
      default: o = v * alpha; break;
    }
    y[i] = o;
  }
}
LOTUS_API int lotus_pointwise_f32(const float* x, const float* a, float* y, int64_t n, int op, float alpha, float drop_p,
                                  uint64_t seed, uint32_t drop_stream, void* stream_) {
  if (n <= 0) return 0;
  k_pointwise<<<min(ceil_div(n, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>(x, a, y, n, op, alpha, drop_p, seed, drop_stream);
  LOTUS_CUDA_OK();
  return 0;
}

// dropout: y = keep(seed, stream_id, i) ? x / (1-p) : 0 — the same call with dy gives the backward
__global__ void k_dropout(const float* __restrict__ x, float* __restrict__ y, int64_t n, float p, uint64_t seed,
                          uint32_t stream_id) {
  float sc = 1.f / (1.f - p);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = dropout_keep(seed, stream_id ^ (uint32_t)(i >> 32), (uint32_t)i, p) ? x[i] * sc : 0.f;
}
LOTUS_API int lotus_dropout_f32(const float* x, float* y, int64_t n, float p, uint64_t seed, uint32_t stream_id,
                                void* stream_) {
  if (n <= 0) return 0;
  LOTUS_CHECK_ARG(p >= 0.f && p < 1.f, "lotus_dropout_f32: p=%f", p);
  k_dropout<<<min(ceil_div(n, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>(x, y, n, p, seed, stream_id);
  LOTUS_CUDA_OK();
  return 0;
}
