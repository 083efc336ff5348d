// Training-step tail on ONE flat fp32 buffer (SURVEY §5 "Distributed communication backend", §8f.3):
// global grad-norm (clip_grad_norm_, train_simple_policy.py:237-241) and the reference's HF-style AdamW
// (train/optim/adamw.py:53-112: eps added after sqrt, bias correction folded into the step size, decoupled
// decay `p -= lr*wd*p` applied after the Adam update with the scheduled lr).
#include "common.cuh"

__global__ void k_sumsq(const float* __restrict__ g, int64_t n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = g[i];
    s = fmaf(v, v, s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
// out_sumsq: one float, zeroed here.
LOTUS_API int lotus_grad_sumsq_f32(const float* grad, int64_t n, float* out_sumsq, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  LOTUS_CUDA_CALL(cudaMemsetAsync(out_sumsq, 0, 4, st));
  if (n > 0) k_sumsq<<<min(ceil_div(n, 1024), kNumSM * 4), 256, 0, st>>>(grad, n, out_sumsq);
  LOTUS_CUDA_OK();
  return 0;
}

// decay_mask[i] (uint8 per PARAMETER TENSOR is expanded by the host into segment boundaries): here the flat
// buffer is described by `seg_end` (exclusive prefix ends, int64) and `seg_wd` (weight decay per segment).
__global__ void k_adamw(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                        int64_t n, const int64_t* __restrict__ seg_end, const float* __restrict__ seg_wd, int nseg,
                        float lr, float beta1, float beta2, float eps, float step_size, const float* __restrict__ sumsq,
                        float max_norm, float grad_scale) {
  // clip coefficient exactly as torch.nn.utils.clip_grad_norm_: min(1, max_norm / (norm + 1e-6))
  float coef = grad_scale;
  if (sumsq && max_norm > 0.f) {
    float norm = sqrtf(*sumsq) * grad_scale;
    float c = max_norm / (norm + 1e-6f);
    coef = grad_scale * fminf(c, 1.f);
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // segment lookup: binary search (nseg ~ 421)
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (seg_end[mid] > i) hi = mid; else lo = mid + 1;
    }
    float wd = seg_wd[lo];
    float gi = g[i] * coef;
    float mi = m[i] * beta1 + (1.f - beta1) * gi;
    float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float pi = p[i];
    pi = pi - step_size * (mi / (sqrtf(vi) + eps));
    if (wd > 0.f) pi = pi - lr * wd * pi;
    p[i] = pi;
  }
}
// float4 form: parameter slots start at multiples of 4 elements (the trainer aligns them to 32), so a vector never
// straddles two segments and the segment search runs once per 4 elements
__global__ void __launch_bounds__(256) k_adamw_v4(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                  float4* __restrict__ v, int64_t n4, const int64_t* __restrict__ seg_end,
                                                  const float* __restrict__ seg_wd, int nseg, float lr, float beta1, float beta2,
                                                  float eps, float step_size, const float* __restrict__ sumsq, float max_norm,
                                                  float grad_scale) {
  float coef = grad_scale;
  if (sumsq && max_norm > 0.f) {
    float norm = sqrtf(*sumsq) * grad_scale;
    float c = max_norm / (norm + 1e-6f);
    coef = grad_scale * fminf(c, 1.f);
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = nseg - 1;
    const int64_t e = i * 4;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (seg_end[mid] > e) hi = mid; else lo = mid + 1;
    }
    const float wd = seg_wd[lo];
    float4 gi = g[i], mi = m[i], vi = v[i], pi = p[i];
    float* gp = (float*)&gi; float* mp = (float*)&mi; float* vp = (float*)&vi; float* pp = (float*)&pi;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gg = gp[k] * coef;
      mp[k] = mp[k] * beta1 + (1.f - beta1) * gg;
      vp[k] = vp[k] * beta2 + (1.f - beta2) * gg * gg;
      float x = pp[k] - step_size * (mp[k] / (sqrtf(vp[k]) + eps));
      if (wd > 0.f) x = x - lr * wd * x;
      pp[k] = x;
    }
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}
LOTUS_API int lotus_adamw_step_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                   const int64_t* seg_end, const float* seg_wd, int nseg, float lr, float beta1,
                                   float beta2, float eps, int step, int flags, const float* grad_sumsq,
                                   float max_grad_norm, float grad_scale, void* stream_) {
  if (n <= 0) return 0;
  double ss = lr;
  if (flags & 1) {
    double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    ss = lr * sqrt(bc2) / bc1;
  }
  // flags: bit 0 = bias correction (the reference's `correct_bias`); bit 1 = every segment boundary is a multiple of 4
  // elements (the trainer pads slots to 32), which enables the float4 path when the buffers are 16-byte aligned
  auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  if ((flags & 2) && n % 4 == 0 && al16(param) && al16(grad) && al16(exp_avg) && al16(exp_avg_sq)) {
    k_adamw_v4<<<min(ceil_div(n / 4, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>(
        (float4*)param, (const float4*)grad, (float4*)exp_avg, (float4*)exp_avg_sq, n / 4, seg_end, seg_wd, nseg, lr, beta1, beta2, eps,
        (float)ss, grad_sumsq, max_grad_norm, grad_scale);
    LOTUS_CUDA_OK();
    return 0;
  }
  k_adamw<<<min(ceil_div(n, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>(
      param, grad, exp_avg, exp_avg_sq, n, seg_end, seg_wd, nseg, lr, beta1, beta2, eps, (float)ss, grad_sumsq,
      max_grad_norm, grad_scale);
  LOTUS_CUDA_OK();
  return 0;
}
