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

// ---------------------------------------------------------------------------------------------- inference front end: crop + voxel down-sample
// reference: evaluation/eval_simple_policy.py:151-175 (Actioner.process_point_clouds): keep the points inside the robot workspace
// (and above the table), then open3d 0.18.0 `PointCloud.voxel_down_sample_and_trace(voxel_size, min(xyz), max(xyz))` (third party, not
// under /root/reference; requirements.txt).  Its published algorithm (open3d/geometry/PointCloud.cpp, VoxelDownSampleAndTrace):
//   voxel_min_bound = min_bound - voxel_size / 2 ;  voxel index = floor((p - voxel_min_bound) / voxel_size)   (float64)
//   output point of a voxel = mean of its points ;  trace[v] = indices of its points in input order
// and the call site keeps `rgb[trace[v][0]]`: the colour of the FIRST input point of the voxel.  open3d emits voxels in the
// iteration order of an unordered_map; here the order is fixed: voxels appear in the order of their first input point.
//
// Three stream-ordered passes over the H x W x cameras points (no sort): (1) crop + 64-bit voxel key + open-addressing hash insert with
// float64 sums, counts and atomicMin of the point index; (2) flag the first point of every voxel and scan the flags; (3) write
// mean / colour / first index at the scanned position.  meta[0] = number of voxels, meta[1] = points inside the crop.
struct VoxSlot { unsigned long long key; double sx, sy, sz; int cnt, first; };
__global__ void k_vox_init(VoxSlot* __restrict__ tab, int64_t cap, int32_t* __restrict__ meta) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cap; i += (int64_t)gridDim.x * blockDim.x) {
    tab[i].key = ~0ull; tab[i].sx = 0; tab[i].sy = 0; tab[i].sz = 0; tab[i].cnt = 0; tab[i].first = 0x7fffffff;
  }
  if (blockIdx.x == 0 && threadIdx.x < 4) meta[threadIdx.x] = 0;
}
__global__ void k_vox_insert(const float* __restrict__ xyz, int n, const double* __restrict__ bounds /* crop lo[3], hi[3], voxel_min_bound[3] */,
                             double voxel_size, VoxSlot* __restrict__ tab, uint64_t cap_mask, int32_t* __restrict__ slot_of_point,
                             int32_t* __restrict__ meta) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double x = xyz[(int64_t)i * 3], y = xyz[(int64_t)i * 3 + 1], z = xyz[(int64_t)i * 3 + 2];
    const bool in = x > bounds[0] && x < bounds[3] && y > bounds[1] && y < bounds[4] && z > bounds[2] && z < bounds[5];
    int slot = -1;
    if (in) {
      const long long vx = (long long)floor((x - bounds[6]) / voxel_size), vy = (long long)floor((y - bounds[7]) / voxel_size),
                      vz = (long long)floor((z - bounds[8]) / voxel_size);
      const unsigned long long key = ((unsigned long long)(vx & 0x1FFFFF) << 42) | ((unsigned long long)(vy & 0x1FFFFF) << 21) | (unsigned long long)(vz & 0x1FFFFF);
      uint64_t k2 = key;
      k2 ^= k2 >> 33; k2 *= 0xff51afd7ed558ccdull; k2 ^= k2 >> 33; k2 *= 0xc4ceb9fe1a85ec53ull; k2 ^= k2 >> 33;
      uint64_t h = k2 & cap_mask;
      while (true) {
        unsigned long long prev = atomicCAS(&tab[h].key, ~0ull, key);
        if (prev == ~0ull || prev == key) break;
        h = (h + 1) & cap_mask;
      }
      slot = (int)h;
      atomicAdd(&tab[h].sx, x); atomicAdd(&tab[h].sy, y); atomicAdd(&tab[h].sz, z);
      atomicAdd(&tab[h].cnt, 1);
      atomicMin(&tab[h].first, i);
      atomicAdd(&meta[1], 1);
    }
    slot_of_point[i] = slot;
  }
}
// single block: flags (first point of its voxel) -> exclusive scan -> position ; meta[0] = voxel count
__global__ void __launch_bounds__(1024) k_vox_scan(const VoxSlot* __restrict__ tab, const int32_t* __restrict__ slot_of_point, int n,
                                                   int32_t* __restrict__ pos, int32_t* __restrict__ meta) {
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    int f = 0;
    if (i < n) { int s = slot_of_point[i]; f = (s >= 0 && tab[s].first == i) ? 1 : 0; }
    int v = f;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) wsum[w] = v;
    __syncthreads();
    if (w == 0) {
      int t = wsum[lane];
      for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += u; }
      wsum[lane] = t;
    }
    __syncthreads();
    const int before = carry + (w > 0 ? wsum[w - 1] : 0) + v - f;
    if (i < n) pos[i] = f ? before : -1;
    __syncthreads();
    if (threadIdx.x == 1023) carry = carry + wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) meta[0] = carry;
}
__global__ void k_vox_write(const VoxSlot* __restrict__ tab, const int32_t* __restrict__ slot_of_point, const int32_t* __restrict__ pos,
                            const uint8_t* __restrict__ rgb, int n, double* __restrict__ out_xyz, uint8_t* __restrict__ out_rgb,
                            int32_t* __restrict__ out_first) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int p = pos[i];
    if (p < 0) continue;
    const VoxSlot& s = tab[slot_of_point[i]];
    const double c = (double)s.cnt;
    out_xyz[(int64_t)p * 3] = s.sx / c; out_xyz[(int64_t)p * 3 + 1] = s.sy / c; out_xyz[(int64_t)p * 3 + 2] = s.sz / c;
    if (rgb) { out_rgb[(int64_t)p * 3] = rgb[(int64_t)i * 3]; out_rgb[(int64_t)p * 3 + 1] = rgb[(int64_t)i * 3 + 1]; out_rgb[(int64_t)p * 3 + 2] = rgb[(int64_t)i * 3 + 2]; }
    out_first[p] = i;
  }
}

LOTUS_API int64_t lotus_voxel_downsample_scratch_bytes(int n) {
  int64_t cap = 1024;
  while (cap < 2 * (int64_t)n) cap <<= 1;
  return cap * (int64_t)sizeof(VoxSlot) + 2 * (int64_t)n * 4 + 64;
}
// xyz [n,3] f32 (all cameras, row-major), rgb [n,3] u8 or null, bounds (device, 9 doubles): workspace crop lo / hi (strict inequalities;
// fold the table height into lo[2]) and open3d's voxel_min_bound = min(cropped xyz) - voxel_size / 2.  Outputs sized for n voxels:
// out_xyz f64 [.,3], out_rgb u8 [.,3], out_first int32 [.]; meta int32[4]: [0] voxels, [1] points inside the crop.
LOTUS_API int lotus_voxel_downsample(const float* xyz, const uint8_t* rgb, int n, const double* bounds, double voxel_size, double* out_xyz,
                                     uint8_t* out_rgb, int32_t* out_first, int32_t* meta, void* scratch, void* stream_) {
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  int64_t cap = 1024;
  while (cap < 2 * (int64_t)n) cap <<= 1;
  VoxSlot* tab = (VoxSlot*)scratch;
  int32_t* slot_of_point = (int32_t*)(tab + cap);
  int32_t* pos = slot_of_point + n;
  k_vox_init<<<min((int)((cap + 255) / 256), kNumSM * 8), 256, 0, st>>>(tab, cap, meta);
  k_vox_insert<<<min(ceil_div(n, 256), kNumSM * 8), 256, 0, st>>>(xyz, n, bounds, voxel_size, tab, (uint64_t)(cap - 1), slot_of_point, meta);
  k_vox_scan<<<1, 1024, 0, st>>>(tab, slot_of_point, n, pos, meta);
  k_vox_write<<<min(ceil_div(n, 256), kNumSM * 8), 256, 0, st>>>(tab, slot_of_point, pos, rgb, n, out_xyz, out_rgb, out_first);
  LOTUS_CUDA_OK();
  return 0;
}
// min / max over the points inside the crop (float64 [6]: min xyz, max xyz), initialised to +-inf by the caller's memset pattern
__global__ void k_crop_minmax(const float* __restrict__ xyz, int n, const double* __restrict__ crop, unsigned long long* __restrict__ mm) {
  // order-preserving map of a double onto uint64 so that atomicMin / atomicMax work on the bit patterns
  auto enc = [](double v) { unsigned long long b = (unsigned long long)__double_as_longlong(v); return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull); };
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double x = xyz[(int64_t)i * 3], y = xyz[(int64_t)i * 3 + 1], z = xyz[(int64_t)i * 3 + 2];
    if (x > crop[0] && x < crop[3] && y > crop[1] && y < crop[4] && z > crop[2] && z < crop[5]) {
      atomicMin(&mm[0], enc(x)); atomicMin(&mm[1], enc(y)); atomicMin(&mm[2], enc(z));
      atomicMax(&mm[3], enc(x)); atomicMax(&mm[4], enc(y)); atomicMax(&mm[5], enc(z));
    }
  }
}
__global__ void k_crop_bounds_finish(const unsigned long long* __restrict__ mm, const double* __restrict__ crop, double voxel_size, double* __restrict__ bounds) {
  const int t = threadIdx.x;
  auto dec = [](unsigned long long e) { unsigned long long b = (e & 0x8000000000000000ull) ? (e & 0x7FFFFFFFFFFFFFFFull) : ~e; return __longlong_as_double((long long)b); };
  if (t < 6) bounds[t] = crop[t];
  if (t < 3) bounds[6 + t] = dec(mm[t]) - voxel_size * 0.5;
}
// bounds[9] (device) from the crop box: the crop itself and voxel_min_bound = min over the cropped points - voxel_size / 2
// (eval_simple_policy.py:171-173 passes np.min(xyz, 0) of the cropped cloud).  scratch: 6 x uint64.
LOTUS_API int lotus_voxel_bounds(const float* xyz, int n, const double* crop_lo_hi, double voxel_size, double* bounds, void* scratch, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  unsigned long long* mm = (unsigned long long*)scratch;
  LOTUS_CUDA_CALL(cudaMemsetAsync(mm, 0xFF, 3 * 8, st));
  LOTUS_CUDA_CALL(cudaMemsetAsync(mm + 3, 0x00, 3 * 8, st));
  if (n > 0) k_crop_minmax<<<min(ceil_div(n, 256), kNumSM * 8), 256, 0, st>>>(xyz, n, crop_lo_hi, mm);
  k_crop_bounds_finish<<<1, 32, 0, st>>>(mm, crop_lo_hi, voxel_size, bounds);
  LOTUS_CUDA_OK();
  return 0;
}
