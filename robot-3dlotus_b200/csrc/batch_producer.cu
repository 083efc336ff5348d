// Device-side batch producer pieces: what the reference's CPU DataLoader workers compute per keystep
// and ship to the GPU.  The soft position labels are 3 * n_i * 2*pos_bins floats per keystep (~1.5 MB, larger than the
// cloud itself); producing them on the device from (xyz, gt position) removes ~93 % of the step's host->device bytes.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------- soft position labels
// reference: genrobo3d/utils/action_position_utils.py:7-46 (get_disc_gt_pos_prob; call site
// train/datasets/simple_policy_dataset.py:346-354).  For cloud b, axis c, point n, bin k:
//   cand = xyz[n,c] + (k - pos_bins) * bin_size ; d = |gt[b,c] - cand|          (float64 like numpy)
//   'dist'  : w = (d > 0.01) ? 0 : 1 / max(d, 1e-4)        'plain': w = (d < 0.01) ? 1 : 0
//   if sum_nk w == 0: w[first argmin d] = 1 ;  prob = w / sum w
// out is the packed concatenation over clouds of (3, n_b * 2*pos_bins) row-major blocks (= the collate's list, flattened).
// scratch: 3*B doubles (sum) + 3*B uint64 (min distance bits) + 3*B int32 (first index of the min), zero / ~0 / INT_MAX
// initialised by the init kernel.
__device__ __forceinline__ double pos_prob_dist(const float* xyz, int64_t ld, int pt, int c, int k, int pos_bins, double bin_size, double g) {
  double cand = (double)(k - pos_bins) * bin_size + (double)xyz[(int64_t)pt * ld + c];
  return fabs(g - cand);
}
__device__ __forceinline__ double pos_prob_weight(double d, int plain) {
  if (plain) return d < 0.01 ? 1.0 : 0.0;
  return d > 0.01 ? 0.0 : 1.0 / fmax(d, 1e-4);
}
__global__ void k_pos_prob_init(double* sum, unsigned long long* dmin, int32_t* imin, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { sum[i] = 0.0; dmin[i] = ~0ull; imin[i] = 0x7fffffff; }
}
__global__ void __launch_bounds__(256) k_pos_prob_stats(const float* __restrict__ xyz, int64_t ld, const int32_t* __restrict__ off,
                                                        const float* __restrict__ gt, int64_t gt_ld, int pos_bins, double bin_size,
                                                        int plain, double* __restrict__ sum, unsigned long long* __restrict__ dmin) {
  __shared__ double rs[8];
  __shared__ unsigned long long rm[8];
  const int b = blockIdx.y, c = blockIdx.z, NB = 2 * pos_bins;
  const int s = off[b], n = off[b + 1] - s;
  const double g = (double)gt[(int64_t)b * gt_ld + c];
  double acc = 0.0;
  unsigned long long m = ~0ull;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    for (int k = 0; k < NB; ++k) {
      double d = pos_prob_dist(xyz, ld, s + i, c, k, pos_bins, bin_size, g);
      acc += pos_prob_weight(d, plain);
      m = min(m, (unsigned long long)__double_as_longlong(d));       // d >= 0: the bit pattern orders like the value
    }
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { rs[w] = acc; rm[w] = m; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { acc += rs[i]; m = min(m, rm[i]); }
    atomicAdd(&sum[b * 3 + c], acc);
    atomicMin(&dmin[b * 3 + c], m);
  }
}
__global__ void __launch_bounds__(256) k_pos_prob_write(const float* __restrict__ xyz, int64_t ld, const int32_t* __restrict__ off,
                                                        const float* __restrict__ gt, int64_t gt_ld, int pos_bins, double bin_size,
                                                        int plain, const double* __restrict__ sum, const unsigned long long* __restrict__ dmin,
                                                        int32_t* __restrict__ imin, float* __restrict__ out) {
  const int b = blockIdx.y, c = blockIdx.z, NB = 2 * pos_bins;
  const int s = off[b], n = off[b + 1] - s;
  const double g = (double)gt[(int64_t)b * gt_ld + c];
  const double tot = sum[b * 3 + c];
  const unsigned long long dm = dmin[b * 3 + c];
  float* o = out + (int64_t)3 * NB * s + (int64_t)c * n * NB;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n * NB; t += gridDim.x * blockDim.x) {
    int i = t / NB, k = t - i * NB;
    double d = pos_prob_dist(xyz, ld, s + i, c, k, pos_bins, bin_size, g);
    if (tot > 0.0) {
      o[t] = (float)(pos_prob_weight(d, plain) / tot);
    } else {
      o[t] = 0.f;                                                      // degenerate axis: one-hot at the first closest candidate
      if ((unsigned long long)__double_as_longlong(d) == dm) atomicMin(&imin[b * 3 + c], t);
    }
  }
}
__global__ void k_pos_prob_fallback(const int32_t* __restrict__ off, int B, int pos_bins, const double* __restrict__ sum,
                                    const int32_t* __restrict__ imin, float* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * B) return;
  int b = t / 3, c = t - 3 * b, NB = 2 * pos_bins;
  int s = off[b], n = off[b + 1] - s;
  if (n > 0 && !(sum[t] > 0.0) && imin[t] != 0x7fffffff) out[(int64_t)3 * NB * s + (int64_t)c * n * NB + imin[t]] = 1.f;
}

// scratch_bytes >= lotus_disc_pos_prob_scratch_bytes(B)
LOTUS_API int64_t lotus_disc_pos_prob_scratch_bytes(int B) { return (int64_t)3 * B * (8 + 8 + 4) + 64; }

LOTUS_API int lotus_disc_pos_prob(const float* xyz, int64_t xyz_ld, const int32_t* cloud_off, int B, int max_cloud, const float* gt_pos,
                                  int64_t gt_ld, int pos_bins, double bin_size, int plain, float* out, void* scratch, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (B <= 0 || max_cloud <= 0) return 0;
  LOTUS_CHECK_ARG(pos_bins > 0 && bin_size > 0, "lotus_disc_pos_prob: pos_bins=%d, bin_size=%g", pos_bins, bin_size);
  LOTUS_CHECK_ARG(((uintptr_t)scratch & 7) == 0, "lotus_disc_pos_prob: scratch must be 8-byte aligned");
  double* sum = (double*)scratch;
  unsigned long long* dmin = (unsigned long long*)(sum + 3 * B);
  int32_t* imin = (int32_t*)(dmin + 3 * B);
  k_pos_prob_init<<<ceil_div(3 * B, 128), 128, 0, st>>>(sum, dmin, imin, 3 * B);
  dim3 g1(min(16, ceil_div(max_cloud, 256)), B, 3);
  k_pos_prob_stats<<<g1, 256, 0, st>>>(xyz, xyz_ld, cloud_off, gt_pos, gt_ld, pos_bins, bin_size, plain, sum, dmin);
  dim3 g2(min(64, ceil_div((int64_t)max_cloud * 2 * pos_bins, 1024)), B, 3);
  k_pos_prob_write<<<g2, 256, 0, st>>>(xyz, xyz_ld, cloud_off, gt_pos, gt_ld, pos_bins, bin_size, plain, sum, dmin, imin, out);
  k_pos_prob_fallback<<<ceil_div(3 * B, 128), 128, 0, st>>>(cloud_off, B, pos_bins, sum, imin, out);
  LOTUS_CUDA_OK();
  return 0;
}
