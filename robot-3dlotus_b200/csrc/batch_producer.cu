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

// ---------------------------------------------------------------------------------------------- geometric augmentation producer
// reference: train/datasets/simple_policy_dataset.py — the per-keystep work of the 4 DataLoader workers between the LMDB read and the
// collate: point subsampling (:281-294), height (:296), z-rotation + U(0, 2 mm) jitter (:158-181 with
// common.py:81-87), centring (:321-336), colour scaling and feature assembly (:342-346).  The robot / table masks (:262-269) and
// the gripper-quaternion rotation + Euler discretisation of the action label (scalar work per keystep, scipy) stay on the host: the
// host draws the rotation angle and hands it over.
//
// (1) lotus_subsample_keysteps: a uniformly random subset of `count[b]` points of raw cloud b in random order ( = the first count[b]
//     entries of np.random.permutation(len), or np.random.choice(len, count, replace=False)): every point gets a counter-based random
//     64-bit key (seed, cloud, index), the cloud's keys are sorted in shared memory (bitonic) and the first count[b] indices are kept.
// (2) lotus_augment_keysteps: for the selected points, in float64 like numpy:
//       h = z - table_height ; p = Rz(angle_b) p ; p += noise (explicit, or U(0, noise_amp) from a counter-based generator) ;
//       centroid_b = mean_b(p) ; p -= centroid_b ; rgb = rgb / 255 * 2 - 1 ; pc_fts = [p, rgb, h] (float32)
//     and gt_pos_out[b] = Rz(angle_b) gt_pos_in[b] - centroid_b.
__device__ __forceinline__ unsigned long long producer_key(uint64_t seed, uint32_t cloud, uint32_t i) {
  uint32_t a = lotus_hash3((uint32_t)seed ^ cloud * 0x27D4EB2Fu, (uint32_t)(seed >> 32), i);
  uint32_t b = lotus_hash3((uint32_t)(seed >> 32) ^ 0x9E3779B1u, (uint32_t)seed + cloud, i ^ 0x5bd1e995u);
  return ((unsigned long long)a << 32) | b;
}
__global__ void k_subsample_sort(const int32_t* __restrict__ raw_off, const int32_t* __restrict__ out_off, uint64_t seed, int np2_max,
                                 int32_t* __restrict__ sel, int32_t* __restrict__ err) {
  extern __shared__ unsigned long long skey[];          // key (high 40 bits of the hash) << 24 | local index
  const int b = blockIdx.x;
  const int s = raw_off[b], n = raw_off[b + 1] - s, cnt = out_off[b + 1] - out_off[b];
  if (n <= 0 || cnt <= 0) return;
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  if (np2 > np2_max || cnt > n || n >= (1 << 24)) { if (threadIdx.x == 0) atomicExch(err, 1); return; }
  for (int i = threadIdx.x; i < np2; i += blockDim.x)
    skey[i] = i < n ? ((producer_key(seed, (uint32_t)b, (uint32_t)i) >> 24) << 24) | (unsigned long long)i : ~0ull;
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long x = skey[i], y = skey[ixj];
          bool up = (i & k) == 0;
          if ((x > y) == up) { skey[i] = y; skey[ixj] = x; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) sel[out_off[b] + i] = s + (int)(skey[i] & 0xFFFFFFull);
}

LOTUS_API int lotus_subsample_keysteps(const int32_t* raw_off, const int32_t* out_off, int B, int max_raw_cloud, uint64_t seed, int32_t* sel_idx,
                                       int32_t* err_flag, void* stream_) {
  if (B <= 0) return 0;
  int np2 = 1;
  while (np2 < max_raw_cloud) np2 <<= 1;
  LOTUS_CHECK_ARG(np2 <= 16384, "lotus_subsample_keysteps: raw clouds of at most 16384 points (got %d)", max_raw_cloud);
  size_t smem = (size_t)np2 * 8;
  static bool attr = false;
  if (!attr) { LOTUS_CUDA_CALL(cudaFuncSetAttribute(k_subsample_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8)); attr = true; }
  LOTUS_CUDA_CALL(cudaMemsetAsync(err_flag, 0, 4, (cudaStream_t)stream_));
  k_subsample_sort<<<B, 512, smem, (cudaStream_t)stream_>>>(raw_off, out_off, seed, np2, sel_idx, err_flag);
  LOTUS_CUDA_OK();
  return 0;
}

__device__ __forceinline__ double producer_uniform(uint64_t seed, uint32_t cloud, uint32_t idx) {    // [0, 1)
  unsigned long long k = producer_key(seed ^ 0xA5A5A5A55A5A5A5Aull, cloud, idx);
  return (double)(k >> 11) * (1.0 / 9007199254740992.0);
}
// pass 1: per-cloud sums of the rotated + jittered coordinates (float64)
__global__ void __launch_bounds__(256) k_augment_sum(const float* __restrict__ raw_xyz, const int32_t* __restrict__ sel, const int32_t* __restrict__ out_off,
                                                     const double* __restrict__ angle, const double* __restrict__ noise, double noise_amp, uint64_t seed,
                                                     double* __restrict__ csum) {
  __shared__ double red[3][8];
  const int b = blockIdx.y;
  const int s = out_off[b], n = out_off[b + 1] - s;
  const double ca = cos(angle[b]), sa = sin(angle[b]);
  double a0 = 0, a1 = 0, a2 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int src = sel ? sel[s + i] : s + i;
    const double x = raw_xyz[(int64_t)src * 3], y = raw_xyz[(int64_t)src * 3 + 1], z = raw_xyz[(int64_t)src * 3 + 2];
    double px = x * ca + y * (-sa), py = x * sa + y * ca, pz = z;                 // np.dot(pc, R^T), R = [[c,-s,0],[s,c,0],[0,0,1]]
    if (noise) { px += noise[(int64_t)(s + i) * 3]; py += noise[(int64_t)(s + i) * 3 + 1]; pz += noise[(int64_t)(s + i) * 3 + 2]; }
    else if (noise_amp > 0) {
      px += noise_amp * producer_uniform(seed, (uint32_t)b, (uint32_t)(3 * i)); py += noise_amp * producer_uniform(seed, (uint32_t)b, (uint32_t)(3 * i + 1));
      pz += noise_amp * producer_uniform(seed, (uint32_t)b, (uint32_t)(3 * i + 2));
    }
    a0 += px; a1 += py; a2 += pz;
  }
  for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][w] = a0; red[1][w] = a1; red[2][w] = a2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += red[threadIdx.x][i];
    atomicAdd(&csum[b * 3 + threadIdx.x], t);
  }
}
// pass 2: features
__global__ void __launch_bounds__(256) k_augment_write(const float* __restrict__ raw_xyz, const uint8_t* __restrict__ raw_rgb, const int32_t* __restrict__ sel,
                                                       const int32_t* __restrict__ out_off, const double* __restrict__ angle, const double* __restrict__ noise,
                                                       double noise_amp, uint64_t seed, double table_height, const double* __restrict__ csum,
                                                       const double* __restrict__ gt_pos_in, int center, float* __restrict__ pc_fts, int fts_ld,
                                                       double* __restrict__ centroid, float* __restrict__ gt_pos_out, int64_t gt_ld) {
  const int b = blockIdx.y;
  const int s = out_off[b], n = out_off[b + 1] - s;
  const double ca = cos(angle[b]), sa = sin(angle[b]);
  const double inv = n > 0 ? 1.0 / (double)n : 0.0;
  const double c0 = center ? csum[b * 3] * inv : 0.0, c1 = center ? csum[b * 3 + 1] * inv : 0.0, c2 = center ? csum[b * 3 + 2] * inv : 0.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    centroid[b * 3] = c0; centroid[b * 3 + 1] = c1; centroid[b * 3 + 2] = c2;
    if (gt_pos_in && gt_pos_out) {
      const double x = gt_pos_in[b * 3], y = gt_pos_in[b * 3 + 1], z = gt_pos_in[b * 3 + 2];
      gt_pos_out[b * gt_ld] = (float)(x * ca + y * (-sa) - c0); gt_pos_out[b * gt_ld + 1] = (float)(x * sa + y * ca - c1); gt_pos_out[b * gt_ld + 2] = (float)(z - c2);
    }
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int src = sel ? sel[s + i] : s + i;
    const double x = raw_xyz[(int64_t)src * 3], y = raw_xyz[(int64_t)src * 3 + 1], z = raw_xyz[(int64_t)src * 3 + 2];
    double px = x * ca + y * (-sa), py = x * sa + y * ca, pz = z;
    if (noise) { px += noise[(int64_t)(s + i) * 3]; py += noise[(int64_t)(s + i) * 3 + 1]; pz += noise[(int64_t)(s + i) * 3 + 2]; }
    else if (noise_amp > 0) {
      px += noise_amp * producer_uniform(seed, (uint32_t)b, (uint32_t)(3 * i)); py += noise_amp * producer_uniform(seed, (uint32_t)b, (uint32_t)(3 * i + 1));
      pz += noise_amp * producer_uniform(seed, (uint32_t)b, (uint32_t)(3 * i + 2));
    }
    float* o = pc_fts + (int64_t)(s + i) * fts_ld;
    o[0] = (float)(px - c0); o[1] = (float)(py - c1); o[2] = (float)(pz - c2);
    int k = 3;
    if (raw_rgb) {
      o[3] = (float)((double)raw_rgb[(int64_t)src * 3] / 255.0 * 2.0 - 1.0);
      o[4] = (float)((double)raw_rgb[(int64_t)src * 3 + 1] / 255.0 * 2.0 - 1.0);
      o[5] = (float)((double)raw_rgb[(int64_t)src * 3 + 2] / 255.0 * 2.0 - 1.0);
      k = 6;
    }
    // height is taken BEFORE the augmentation (:296), in float32: numpy keeps `float32 array - python float` in float32
    if (fts_ld > k) o[k] = raw_xyz[(int64_t)src * 3 + 2] - (float)table_height;
  }
}

// raw_xyz [M,3] f32, raw_rgb [M,3] u8 (or null: xyz + height features, the motion planner's `use_color False`), sel_idx [N] (or null:
// identity), out_off [B+1] offsets of the output clouds, angle [B] f64 radians, noise [N,3] f64 (or null: device generator with
// amplitude noise_amp), gt_pos_in [B,3] f64 (or null).  pc_fts [N, fts_ld] f32 (fts_ld = 7 with colour, 4 without), centroid [B,3] f64,
// gt_pos_out rows of gt_ld floats (the first three are written).  scratch: 3*B doubles.
LOTUS_API int lotus_augment_keysteps(const float* raw_xyz, const uint8_t* raw_rgb, const int32_t* sel_idx, const int32_t* out_off, int B,
                                     int max_cloud, const double* angle, const double* noise, double noise_amp, uint64_t seed,
                                     double table_height, int center, const double* gt_pos_in, float* pc_fts, int fts_ld, double* centroid,
                                     float* gt_pos_out, int64_t gt_ld, double* scratch, void* stream_) {
  if (B <= 0) return 0;
  LOTUS_CHECK_ARG(fts_ld >= (raw_rgb ? 6 : 3), "lotus_augment_keysteps: fts_ld=%d too small", fts_ld);
  cudaStream_t st = (cudaStream_t)stream_;
  LOTUS_CUDA_CALL(cudaMemsetAsync(scratch, 0, (size_t)3 * B * 8, st));
  dim3 grid(max(1, min(ceil_div(max_cloud, 256), 32)), B);
  if (center) k_augment_sum<<<grid, 256, 0, st>>>(raw_xyz, sel_idx, out_off, angle, noise, noise_amp, seed, scratch);
  k_augment_write<<<grid, 256, 0, st>>>(raw_xyz, raw_rgb, sel_idx, out_off, angle, noise, noise_amp, seed, table_height, scratch, gt_pos_in, center,
                                        pc_fts, fts_ld, centroid, gt_pos_out, gt_ld);
  LOTUS_CUDA_OK();
  return 0;
}
