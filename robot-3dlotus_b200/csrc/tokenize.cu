// Integer tokeniser of the 3D-LOTUS hot path (SURVEY §8a rows 1-3, 6, 10 integer half, 12-13).
//
// Everything here is HBM-bound integer work: one thread per point with coalesced SoA/AoS reads, per-cloud
// work staged in shared memory (the segmented sort, the cluster scan, FPS), no host synchronisation —
// data-dependent sizes live in a device-side `meta` block and kernels are launched over host-known upper
// bounds.  Reference behaviour restated (paths relative to /root/reference/genrobo3d/models/PointTransformerV3):
//   grid coords / depth ......... model.py:96-103
//   z-order / Hilbert codes ..... serialization/default.py:10-27, z_order.py:40-50, hilbert.py:143-198
//   argsort / inverse / shuffle . model.py:121-134
//   patch plan .................. model.py:410-466
//   grid-pool clusters .......... model.py:726-771
//   sparse-conv neighbourhoods .. spconv 2.3.6 SubMConv3d indice generation (third party, restated)
//   FPS ......................... ../../train/datasets/common.py:108-129
#include "common.cuh"

// meta block layout (int32 x 8 per stage)
enum { META_N = 0, META_DEPTH = 1, META_NPAD = 2, META_NPATCH = 3, META_MAXCLOUD = 4, META_STRIDE = 8 };

// ------------------------------------------------------------------------------------------------
// grid coordinates
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_coord_min(const float* __restrict__ coord, int64_t n, uint32_t* __restrict__ mn /*3, init 0xFFFFFFFF*/) {
  float m0 = INFINITY, m1 = INFINITY, m2 = INFINITY;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    m0 = fminf(m0, coord[i * 3 + 0]);
    m1 = fminf(m1, coord[i * 3 + 1]);
    m2 = fminf(m2, coord[i * 3 + 2]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m0 = fminf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
    m1 = fminf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
    m2 = fminf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&mn[0], f2ord(m0));
    atomicMin(&mn[1], f2ord(m1));
    atomicMin(&mn[2], f2ord(m2));
  }
}

// mode 0: trunc((c - min) / gs)   (torch CPU kernel: true IEEE division)
// mode 1: trunc((c - min) * (1/gs)) (torch CUDA kernel multiplies by the reciprocal of a scalar divisor)
__global__ void k_grid(const float* __restrict__ coord, int64_t n, const uint32_t* __restrict__ mn, float gs, int mode,
                       int32_t* __restrict__ grid, int32_t* __restrict__ gmax /*init 0*/) {
  float m[3] = {ord2f(mn[0]), ord2f(mn[1]), ord2f(mn[2])};
  float inv = __fdiv_rn(1.0f, gs);
  int mx = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n * 3; i += (int64_t)gridDim.x * blockDim.x) {
    float d = __fsub_rn(coord[i], m[i % 3]);
    float q = mode ? __fmul_rn(d, inv) : __fdiv_rn(d, gs);
    int g = (int)truncf(q);
    grid[i] = g;
    mx = max(mx, g);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(gmax, mx);
}

__global__ void k_finish_meta0(const int32_t* gmax, int32_t n, int32_t* meta) {
  int m = *gmax;
  meta[META_N] = n;
  meta[META_DEPTH] = (m <= 0) ? 0 : (32 - __clz(m));
}

LOTUS_API int lotus_grid_coord(const float* coord, int64_t n, float grid_size, int mode, int32_t* grid, int32_t* meta,
                               void* scratch16, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  LOTUS_CHECK_ARG(n > 0 && n < (1ll << 31) / 3, "lotus_grid_coord: n=%lld out of range", (long long)n);
  uint32_t* mn = (uint32_t*)scratch16;
  int32_t* gmax = (int32_t*)scratch16 + 3;
  LOTUS_CUDA_CALL(cudaMemsetAsync(mn, 0xFF, 12, st));
  LOTUS_CUDA_CALL(cudaMemsetAsync(gmax, 0, 4, st));
  int blocks = min(ceil_div(n, 256), kNumSM * 8);
  k_coord_min<<<blocks, 256, 0, st>>>(coord, n, mn);
  k_grid<<<min(ceil_div(n * 3, 256), kNumSM * 8), 256, 0, st>>>(coord, n, mn, grid_size, mode, grid, gmax);
  k_finish_meta0<<<1, 1, 0, st>>>(gmax, (int32_t)n, meta);
  LOTUS_CUDA_OK();
  return 0;
}

// Shape-static form for CUDA-graph replay: the launch covers `n_upper` points, the ACTUAL count is read on the device (n_dev, e.g.
// &cloud_off[B]) so that one captured graph serves every cloud of a size bucket ("CUDA-graph capture with
// shape bucketing").  Points beyond *n_dev are never read.
__global__ void k_coord_min_n(const float* __restrict__ coord, const int32_t* __restrict__ n_dev, uint32_t* __restrict__ mn) {
  const int64_t n = *n_dev;
  float m0 = INFINITY, m1 = INFINITY, m2 = INFINITY;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    m0 = fminf(m0, coord[i * 3 + 0]); m1 = fminf(m1, coord[i * 3 + 1]); m2 = fminf(m2, coord[i * 3 + 2]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m0 = fminf(m0, __shfl_xor_sync(0xffffffffu, m0, o)); m1 = fminf(m1, __shfl_xor_sync(0xffffffffu, m1, o)); m2 = fminf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&mn[0], f2ord(m0)); atomicMin(&mn[1], f2ord(m1)); atomicMin(&mn[2], f2ord(m2)); }
}
__global__ void k_grid_n(const float* __restrict__ coord, const int32_t* __restrict__ n_dev, const uint32_t* __restrict__ mn, float gs, int mode,
                         int32_t* __restrict__ grid, int32_t* __restrict__ gmax) {
  const int64_t n = *n_dev;
  float m[3] = {ord2f(mn[0]), ord2f(mn[1]), ord2f(mn[2])};
  float inv = __fdiv_rn(1.0f, gs);
  int mx = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n * 3; i += (int64_t)gridDim.x * blockDim.x) {
    float d = __fsub_rn(coord[i], m[i % 3]);
    float q = mode ? __fmul_rn(d, inv) : __fdiv_rn(d, gs);
    int g = (int)truncf(q);
    grid[i] = g;
    mx = max(mx, g);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(gmax, mx);
}
__global__ void k_finish_meta0_n(const int32_t* gmax, const int32_t* n_dev, int32_t* meta) {
  int m = *gmax;
  meta[META_N] = *n_dev;
  meta[META_DEPTH] = (m <= 0) ? 0 : (32 - __clz(m));
}
LOTUS_API int lotus_grid_coord_n(const float* coord, int64_t n_upper, const int32_t* n_dev, float grid_size, int mode, int32_t* grid,
                                 int32_t* meta, void* scratch16, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  LOTUS_CHECK_ARG(n_upper > 0 && n_upper < (1ll << 31) / 3, "lotus_grid_coord_n: n_upper=%lld out of range", (long long)n_upper);
  uint32_t* mn = (uint32_t*)scratch16;
  int32_t* gmax = (int32_t*)scratch16 + 3;
  LOTUS_CUDA_CALL(cudaMemsetAsync(mn, 0xFF, 12, st));
  LOTUS_CUDA_CALL(cudaMemsetAsync(gmax, 0, 4, st));
  k_coord_min_n<<<min(ceil_div(n_upper, 256), kNumSM * 8), 256, 0, st>>>(coord, n_dev, mn);
  k_grid_n<<<min(ceil_div(n_upper * 3, 256), kNumSM * 8), 256, 0, st>>>(coord, n_dev, mn, grid_size, mode, grid, gmax);
  k_finish_meta0_n<<<1, 1, 0, st>>>(gmax, n_dev, meta);
  LOTUS_CUDA_OK();
  return 0;
}

// Makes the rows between a stage's ACTUAL point count (device meta) and the static upper bound the kernels are launched over inert:
// neighbour-table rows -> -1 (no neighbour), the pooling cluster of such a row -> 0, and the child stage's CSR pointers beyond its
// actual size -> empty segments.  With that, every kernel of a forward may run over the bucket's row bound: rows past the real count
// compute garbage that no real row ever reads (graph replay with shape bucketing; inference).
__global__ void k_stage_sanitize(const int32_t* __restrict__ meta, int n_upper, int32_t* __restrict__ nbr3, int64_t stride, int32_t* __restrict__ nbr5,
                                 int32_t* __restrict__ cluster, const int32_t* __restrict__ meta_child, int n_child_upper, int32_t* __restrict__ seg_ptr_child) {
  const int n = meta[META_N];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  const int tail = max(0, n_upper - n);
  for (int64_t t = tid; t < (int64_t)tail * 27; t += nt) { int d = (int)(t / tail), m = n + (int)(t % tail); nbr3[(int64_t)d * stride + m] = -1; }
  if (nbr5) for (int64_t t = tid; t < (int64_t)tail * 125; t += nt) { int d = (int)(t / tail), m = n + (int)(t % tail); nbr5[(int64_t)d * stride + m] = -1; }
  if (cluster) for (int t = tid; t < tail; t += nt) cluster[n + t] = 0;
  if (seg_ptr_child) {
    const int nc = meta_child[META_N];
    const int last = seg_ptr_child[nc];
    for (int t = nc + 1 + tid; t <= n_child_upper; t += nt) seg_ptr_child[t] = last;
  }
}
LOTUS_API int lotus_stage_sanitize(const int32_t* meta, int n_upper, int32_t* nbr3, int64_t nbr_stride, int32_t* nbr5, int32_t* cluster,
                                   const int32_t* meta_child, int n_child_upper, int32_t* seg_ptr_child, void* stream_) {
  k_stage_sanitize<<<kNumSM, 256, 0, (cudaStream_t)stream_>>>(meta, n_upper, nbr3, nbr_stride, nbr5, cluster, meta_child, n_child_upper, seg_ptr_child);
  LOTUS_CUDA_OK();
  return 0;
}

// batch id per point from device offsets (model.py:37-42)
__global__ void k_batch_ids(const int32_t* __restrict__ off, int B, int32_t* __restrict__ batch) {
  int b = blockIdx.x;
  for (int i = off[b] + threadIdx.x; i < off[b + 1]; i += blockDim.x) batch[i] = b;
}
LOTUS_API int lotus_batch_ids(const int32_t* cloud_off, int B, int32_t* batch, void* stream_) {
  k_batch_ids<<<B, 256, 0, (cudaStream_t)stream_>>>(cloud_off, B, batch);
  LOTUS_CUDA_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// curve codes
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t spread3(uint32_t v) {  // 16 bits -> every third bit of 48
  uint64_t x = v & 0xFFFFu;
  x = (x | (x << 32)) & 0x00FF00000000FFFFull;  // not used for <=16 bits but keeps the classic ladder shape
  x = (x | (x << 16)) & 0x00FF0000FF0000FFull;
  x = (x | (x << 8)) & 0xF00F00F00F00F00Full;
  x = (x | (x << 4)) & 0x30C30C30C30C30C3ull;
  x = (x | (x << 2)) & 0x9249249249249249ull;
  return x;
}
__device__ __forceinline__ uint64_t interleave3(uint32_t x, uint32_t y, uint32_t z) {
  return (spread3(x) << 2) | (spread3(y) << 1) | spread3(z);
}
__device__ __forceinline__ uint64_t hilbert3(uint32_t x0, uint32_t x1, uint32_t x2, int depth) {
  // Skilling's transform, planes from the most significant down (hilbert.py:156-175)
  uint32_t X[3] = {x0, x1, x2};
  for (int bit = depth - 1; bit >= 0; --bit) {
    uint32_t q = 1u << bit, low = q - 1u;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (X[d] & q) {
        X[0] ^= low;
      } else {
        uint32_t t = (X[0] ^ X[d]) & low;
        X[0] ^= t;
        X[d] ^= t;
      }
    }
  }
  uint64_t g = interleave3(X[0], X[1], X[2]);  // hilbert.py:178
  g ^= g >> 1; g ^= g >> 2; g ^= g >> 4; g ^= g >> 8; g ^= g >> 16; g ^= g >> 32;  // gray2binary, hilbert.py:181
  return g;
}

struct Curves { int c[4]; };  // 0 z, 1 z-trans, 2 hilbert, 3 hilbert-trans

__global__ void k_codes(const int32_t* __restrict__ grid, const int32_t* __restrict__ batch,
                        const int32_t* __restrict__ meta, Curves cv, const int32_t* __restrict__ cv_dev, int nrows, int64_t row_stride,
                        int64_t* __restrict__ code) {
  int n = meta[META_N], depth = meta[META_DEPTH];
  if (cv_dev)                                  // graph replay: the shuffled curve of every row comes from device memory
    for (int r = 0; r < nrows; ++r) cv.c[r] = cv_dev[r] & 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t x = grid[i * 3], y = grid[i * 3 + 1], z = grid[i * 3 + 2];
    uint64_t hi = (uint64_t)batch[i] << (3 * depth);
    uint64_t zc = 0, zt = 0, hc = 0, ht = 0;
    bool need[4] = {false, false, false, false};
    for (int r = 0; r < nrows; ++r) need[cv.c[r]] = true;
    if (need[0]) zc = interleave3(x, y, z);
    if (need[1]) zt = interleave3(y, x, z);
    if (need[2]) hc = hilbert3(x, y, z, depth);
    if (need[3]) ht = hilbert3(y, x, z, depth);
    for (int r = 0; r < nrows; ++r) {
      int c = cv.c[r];
      uint64_t v = c == 0 ? zc : c == 1 ? zt : c == 2 ? hc : ht;
      code[r * row_stride + i] = (int64_t)(hi | v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// per-cloud segmented sort in shared memory (bitonic on key = code_low << idx_bits | local index)
// batch id occupies the high bits of every code, so the global argsort is the concatenation of per-cloud
// sorts; ties (duplicate voxels) resolve by original index (stable).
// ------------------------------------------------------------------------------------------------
__global__ void k_cloud_sort(const int64_t* __restrict__ code, const int32_t* __restrict__ cloud_off,
                             const int32_t* __restrict__ meta, int64_t row_stride, int idx_bits, int npow2_max,
                             int32_t* __restrict__ order, int32_t* __restrict__ inverse, int32_t* __restrict__ err) {
  extern __shared__ uint64_t skeys[];
  int b = blockIdx.x, r = blockIdx.y;
  int s = cloud_off[b], e = cloud_off[b + 1], n = e - s;
  if (n <= 0) return;
  int depth = meta[META_DEPTH];
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  if (np2 > npow2_max || 3 * depth + idx_bits > 64) {
    if (threadIdx.x == 0) atomicExch(err, 1);
    return;
  }
  uint64_t cmask = (depth == 0) ? 0ull : ((3 * depth >= 64) ? ~0ull : ((1ull << (3 * depth)) - 1ull));
  const int64_t* crow = code + r * row_stride;
  for (int i = threadIdx.x; i < np2; i += blockDim.x)
    skeys[i] = (i < n) ? ((((uint64_t)crow[s + i] & cmask) << idx_bits) | (uint64_t)i) : ~0ull;
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j cleared
        int p = i | j;
        uint64_t a = skeys[i], c = skeys[p];
        bool up = ((i & k) == 0);
        if ((a > c) == up) {
          skeys[i] = c;
          skeys[p] = a;
        }
      }
      __syncthreads();
    }
  }
  uint64_t imask = (1ull << idx_bits) - 1ull;
  int32_t* orow = order + r * row_stride;
  int32_t* irow = inverse + r * row_stride;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int src = (int)(skeys[i] & imask);
    orow[s + i] = s + src;
    irow[s + src] = s + i;
  }
}

static int serialize_impl(const int32_t* grid, const int32_t* batch, const int32_t* cloud_off, int B,
                              int n_upper, int max_cloud_upper, const int32_t* meta, const int32_t* curve_of_row,
                              int nrows, int64_t row_stride, int64_t* code, int32_t* order, int32_t* inverse,
                              int32_t* err_flag, const int32_t* curves_dev, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  LOTUS_CHECK_ARG(nrows >= 1 && nrows <= 4, "lotus_serialize: nrows=%d", nrows);
  LOTUS_CHECK_ARG(max_cloud_upper <= 16384, "lotus_serialize: clouds above 16384 points (%d) exceed the shared-memory sort", max_cloud_upper);
  Curves cv;
  for (int r = 0; r < 4; ++r) {
    cv.c[r] = (r < nrows && curve_of_row) ? curve_of_row[r] : 0;
    LOTUS_CHECK_ARG(cv.c[r] >= 0 && cv.c[r] < 4, "lotus_serialize: bad curve id");
  }
  k_codes<<<min(ceil_div(n_upper, 256), kNumSM * 8), 256, 0, st>>>(grid, batch, meta, cv, curves_dev, nrows, row_stride, code);
  int np2 = 1, bits = 0;
  while (np2 < max_cloud_upper) { np2 <<= 1; ++bits; }
  if (bits < 1) bits = 1;
  size_t smem = (size_t)np2 * 8;
  static bool attr_set = false;
  if (!attr_set) {
    LOTUS_CUDA_CALL(cudaFuncSetAttribute(k_cloud_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    attr_set = true;
  }
  int threads = np2 >= 2048 ? 1024 : (np2 >= 512 ? 256 : 128);
  k_cloud_sort<<<dim3(B, nrows), threads, smem, st>>>(code, cloud_off, meta, row_stride, bits, np2, order, inverse, err_flag);
  LOTUS_CUDA_OK();
  return 0;
}

LOTUS_API int lotus_serialize(const int32_t* grid, const int32_t* batch, const int32_t* cloud_off, int B,
                              int n_upper, int max_cloud_upper, const int32_t* meta, const int32_t* curve_of_row,
                              int nrows, int64_t row_stride, int64_t* code, int32_t* order, int32_t* inverse,
                              int32_t* err_flag, void* stream_) {
  return serialize_impl(grid, batch, cloud_off, B, n_upper, max_cloud_upper, meta, curve_of_row, nrows, row_stride, code, order, inverse, err_flag,
                        nullptr, stream_);
}
// curve ids of the rows read from DEVICE memory (curve_of_row_dev[nrows]): the per-forward `shuffle_orders` draw (model.py:130-134,
// 750-754) stays a run-time input of a captured CUDA graph
LOTUS_API int lotus_serialize_dev(const int32_t* grid, const int32_t* batch, const int32_t* cloud_off, int B,
                                  int n_upper, int max_cloud_upper, const int32_t* meta, const int32_t* curve_of_row_dev,
                                  int nrows, int64_t row_stride, int64_t* code, int32_t* order, int32_t* inverse,
                                  int32_t* err_flag, void* stream_) {
  return serialize_impl(grid, batch, cloud_off, B, n_upper, max_cloud_upper, meta, nullptr, nrows, row_stride, code, order, inverse, err_flag,
                        curve_of_row_dev, stream_);
}

// ------------------------------------------------------------------------------------------------
// grid-pool cluster structure (integer half of SerializedPooling.forward)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool cluster_head_flag(const int64_t* code0, const int32_t* order0, int j, int s, int shift) {
  if (j == s) return true;
  return (code0[order0[j]] >> shift) != (code0[order0[j - 1]] >> shift);
}

__global__ void k_pool_count(const int64_t* __restrict__ code0, const int32_t* __restrict__ order0,
                             const int32_t* __restrict__ cloud_off, const int32_t* __restrict__ meta_in, int stride_log2,
                             int32_t* __restrict__ cloud_m) {
  __shared__ int red[32];
  int b = blockIdx.x, s = cloud_off[b], e = cloud_off[b + 1];
  int depth = meta_in[META_DEPTH];
  int pd = stride_log2 > depth ? 0 : stride_log2;
  int cnt = 0;
  for (int j = s + threadIdx.x; j < e; j += blockDim.x) cnt += cluster_head_flag(code0, order0, j, s, 3 * pd) ? 1 : 0;
  cnt = warp_sum_i(cnt);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0;
    v = warp_sum_i(v);
    if (threadIdx.x == 0) cloud_m[b] = v;
  }
}

__global__ void k_pool_scan(const int32_t* __restrict__ cloud_m, int B, const int32_t* __restrict__ meta_in,
                            int stride_log2, int32_t* __restrict__ cloud_off_out, int32_t* __restrict__ meta_out) {
  // B is small (<= a few thousand): one thread scans
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0, mx = 0;
    for (int b = 0; b < B; ++b) {
      cloud_off_out[b] = acc;
      acc += cloud_m[b];
      mx = max(mx, cloud_m[b]);
    }
    cloud_off_out[B] = acc;
    int depth = meta_in[META_DEPTH];
    int pd = stride_log2 > depth ? 0 : stride_log2;
    meta_out[META_N] = acc;
    meta_out[META_DEPTH] = depth - pd;
    meta_out[META_MAXCLOUD] = mx;
  }
}

__global__ void k_pool_assign(const int64_t* __restrict__ code0, const int32_t* __restrict__ order0,
                              const int32_t* __restrict__ grid, const int32_t* __restrict__ cloud_off,
                              const int32_t* __restrict__ cloud_off_out, const int32_t* __restrict__ meta_in,
                              int stride_log2, int B, int32_t* __restrict__ cluster, int32_t* __restrict__ head,
                              int32_t* __restrict__ seg_ptr, int32_t* __restrict__ grid_out,
                              int32_t* __restrict__ batch_out) {
  __shared__ int wsum[32];
  __shared__ int carry;
  int b = blockIdx.x, s = cloud_off[b], e = cloud_off[b + 1];
  int depth = meta_in[META_DEPTH];
  int pd = stride_log2 > depth ? 0 : stride_log2;
  int base = cloud_off_out[b];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j0 = s; j0 < e; j0 += blockDim.x) {
    int j = j0 + threadIdx.x;
    int f = (j < e && cluster_head_flag(code0, order0, j, s, 3 * pd)) ? 1 : 0;
    // inclusive scan of f across the block
    int v = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) wsum[w] = v;
    __syncthreads();
    if (w == 0) {
      int t = lane < nw ? wsum[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u;
      }
      wsum[lane] = t;  // inclusive over warps
    }
    __syncthreads();
    int incl = v + (w > 0 ? wsum[w - 1] : 0) + carry;
    if (j < e) {
      int c = base + incl - 1;
      int p = order0[j];
      cluster[p] = c;
      if (f) {
        head[c] = p;
        seg_ptr[c] = j;
        grid_out[c * 3 + 0] = grid[p * 3 + 0] >> pd;
        grid_out[c * 3 + 1] = grid[p * 3 + 1] >> pd;
        grid_out[c * 3 + 2] = grid[p * 3 + 2] >> pd;
        batch_out[c] = b;
      }
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = incl;
    __syncthreads();
  }
  if (b == B - 1 && threadIdx.x == 0) seg_ptr[cloud_off_out[B]] = e;
}

LOTUS_API int lotus_pool_structure(const int64_t* code0, const int32_t* order0, const int32_t* grid,
                                   const int32_t* cloud_off, int B, int stride, const int32_t* meta_in,
                                   int32_t* cluster, int32_t* head, int32_t* seg_ptr, int32_t* grid_out,
                                   int32_t* batch_out, int32_t* cloud_off_out, int32_t* meta_out, int32_t* scratch_B,
                                   void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  int sl = 0;
  while ((1 << sl) < stride) ++sl;
  LOTUS_CHECK_ARG((1 << sl) == stride, "lotus_pool_structure: stride %d is not a power of two (model.py:699)", stride);
  k_pool_count<<<B, 256, 0, st>>>(code0, order0, cloud_off, meta_in, sl, scratch_B);
  k_pool_scan<<<1, 32, 0, st>>>(scratch_B, B, meta_in, sl, cloud_off_out, meta_out);
  k_pool_assign<<<B, 256, 0, st>>>(code0, order0, grid, cloud_off, cloud_off_out, meta_in, sl, B, cluster, head, seg_ptr,
                                   grid_out, batch_out);
  LOTUS_CUDA_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// patch plan (SerializedAttention.get_padding_and_inverse) fused with order/inverse composition
// ------------------------------------------------------------------------------------------------
__global__ void k_plan_scan(const int32_t* __restrict__ cloud_off, int B, int K, int32_t* __restrict__ pad_off,
                            int32_t* __restrict__ patch_off, int32_t* __restrict__ meta) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int a = 0, p = 0;
    for (int b = 0; b < B; ++b) {
      int c = cloud_off[b + 1] - cloud_off[b];
      int cp = c > K ? ((c + K - 1) / K) * K : c;
      pad_off[b] = a;
      patch_off[b] = p;
      a += cp;
      p += (cp + K - 1) / K;
    }
    pad_off[B] = a;
    patch_off[B] = p;
    meta[META_NPAD] = a;
    meta[META_NPATCH] = p;
  }
}

__global__ void k_plan_fill(const int32_t* __restrict__ cloud_off, const int32_t* __restrict__ pad_off,
                            const int32_t* __restrict__ patch_off, int B, int K, int cu_len,
                            const int32_t* __restrict__ order_row, const int32_t* __restrict__ inverse_row,
                            int32_t* __restrict__ pad, int32_t* __restrict__ unpad, int32_t* __restrict__ gather_idx,
                            int32_t* __restrict__ slot_of_point, int32_t* __restrict__ cu_seqlens) {
  int b = blockIdx.x;
  int s = cloud_off[b], c = cloud_off[b + 1] - s;
  int ps = pad_off[b], cp = pad_off[b + 1] - ps;
  int r = c % K;
  for (int i = threadIdx.x; i < cp; i += blockDim.x) {
    int src = i;
    if (cp != c && i >= cp - K + r) src = i - K;  // tail of the last patch aliases the previous patch's slots
    int sorted_pos = s + src;
    if (pad) pad[ps + i] = sorted_pos;
    gather_idx[ps + i] = order_row[sorted_pos];
  }
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    if (unpad) unpad[s + i] = ps + i;
    // point p sits at sorted position inverse[p]; its slot is unpad[inverse[p]]
    int p = s + i;
    slot_of_point[p] = ps + (inverse_row[p] - s);
  }
  int np = patch_off[b + 1] - patch_off[b];
  for (int i = threadIdx.x; i < np; i += blockDim.x) cu_seqlens[patch_off[b] + i] = ps + i * K;
  if (b == B - 1) {
    int total_p = patch_off[B], total = pad_off[B];
    for (int i = total_p + threadIdx.x; i < cu_len; i += blockDim.x) cu_seqlens[i] = total;
  }
}

LOTUS_API int lotus_patch_plan(const int32_t* cloud_off, int B, int K, const int32_t* order_row,
                               const int32_t* inverse_row, int32_t* pad, int32_t* unpad, int32_t* gather_idx,
                               int32_t* slot_of_point, int32_t* cu_seqlens, int cu_len, int32_t* meta,
                               int32_t* scratch_2B2, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  LOTUS_CHECK_ARG(K > 0 && B > 0, "lotus_patch_plan: bad K/B");
  int32_t* pad_off = scratch_2B2;
  int32_t* patch_off = scratch_2B2 + (B + 1);
  k_plan_scan<<<1, 32, 0, st>>>(cloud_off, B, K, pad_off, patch_off, meta);
  k_plan_fill<<<B, 256, 0, st>>>(cloud_off, pad_off, patch_off, B, K, cu_len, order_row, inverse_row, pad, unpad,
                                 gather_idx, slot_of_point, cu_seqlens);
  LOTUS_CUDA_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// submanifold-conv neighbourhoods: open-addressing hash on (batch,x,y,z), then k^3 probes per voxel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t vox_key(int b, int x, int y, int z) {
  return ((uint64_t)(uint32_t)b << 48) | ((uint64_t)(x & 0xFFFF) << 32) | ((uint64_t)(y & 0xFFFF) << 16) | (uint64_t)(z & 0xFFFF);
}
__device__ __forceinline__ uint64_t mix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}
__global__ void k_hash_insert(const int32_t* __restrict__ grid, const int32_t* __restrict__ batch,
                              const int32_t* __restrict__ meta, unsigned long long* __restrict__ keys,
                              int32_t* __restrict__ vals, uint64_t cap_mask) {
  int n = meta[META_N];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint64_t k = vox_key(batch[i], grid[i * 3], grid[i * 3 + 1], grid[i * 3 + 2]);
    uint64_t h = mix64(k) & cap_mask;
    while (true) {
      unsigned long long prev = atomicCAS(&keys[h], ~0ull, (unsigned long long)k);
      if (prev == ~0ull || prev == k) { atomicMin(&vals[h], i); break; }  // duplicate voxel: lowest index wins
      h = (h + 1) & cap_mask;
    }
  }
}
__global__ void k_hash_lookup(const int32_t* __restrict__ grid, const int32_t* __restrict__ batch,
                              const int32_t* __restrict__ meta, const unsigned long long* __restrict__ keys,
                              const int32_t* __restrict__ vals, uint64_t cap_mask, int ks, int64_t nbr_stride,
                              int32_t* __restrict__ nbr) {
  int n = meta[META_N];
  int k3 = ks * ks * ks, r = ks / 2;
  int64_t total = (int64_t)n * k3;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int d = (int)(t / n), i = (int)(t % n);  // d-major so writes (and later reads) are coalesced over points
    int d0 = d / (ks * ks) - r, d1 = (d / ks) % ks - r, d2 = d % ks - r;
    int x = grid[i * 3] + d0, y = grid[i * 3 + 1] + d1, z = grid[i * 3 + 2] + d2;
    int res = -1;
    if (x >= 0 && y >= 0 && z >= 0 && x < 65536 && y < 65536 && z < 65536) {
      uint64_t k = vox_key(batch[i], x, y, z);
      uint64_t h = mix64(k) & cap_mask;
      while (true) {
        unsigned long long cur = keys[h];
        if (cur == k) { res = vals[h]; break; }
        if (cur == ~0ull) break;
        h = (h + 1) & cap_mask;
      }
    }
    nbr[d * nbr_stride + i] = res;
  }
}

LOTUS_API int lotus_neighbor_table(const int32_t* grid, const int32_t* batch, const int32_t* meta, int n_upper,
                                   int ksize, uint64_t* hash_keys, int32_t* hash_vals, int64_t hash_cap, int build,
                                   int32_t* nbr, int64_t nbr_stride, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  LOTUS_CHECK_ARG(ksize == 3 || ksize == 5, "lotus_neighbor_table: kernel size %d (3 or 5)", ksize);
  LOTUS_CHECK_ARG(hash_cap >= 2 * (int64_t)n_upper && (hash_cap & (hash_cap - 1)) == 0,
                  "lotus_neighbor_table: hash_cap must be a power of two >= 2*n");
  if (build) {
    LOTUS_CUDA_CALL(cudaMemsetAsync(hash_keys, 0xFF, (size_t)hash_cap * 8, st));
    LOTUS_CUDA_CALL(cudaMemsetAsync(hash_vals, 0x7F, (size_t)hash_cap * 4, st));
    k_hash_insert<<<min(ceil_div(n_upper, 256), kNumSM * 8), 256, 0, st>>>(
        grid, batch, meta, (unsigned long long*)hash_keys, hash_vals, (uint64_t)hash_cap - 1);
  }
  int64_t total = (int64_t)n_upper * ksize * ksize * ksize;
  k_hash_lookup<<<min(ceil_div(total, 256), kNumSM * 16), 256, 0, st>>>(
      grid, batch, meta, (const unsigned long long*)hash_keys, hash_vals, (uint64_t)hash_cap - 1, ksize, nbr_stride, nbr);
  LOTUS_CUDA_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// farthest point sampling: one CTA per cloud, xyz staged in shared memory (SoA), running min-distance
// in shared memory, block-wide argmax per iteration (ties -> lowest index, like np.argmax).
// fp32 arithmetic in numpy's order ((dx*dx + dy*dy) + dz*dz), no FMA contraction.
// ------------------------------------------------------------------------------------------------
__global__ void k_fps(const float* __restrict__ xyz, int64_t row_stride, const int32_t* __restrict__ cloud_off,
                      int npoint, const int32_t* __restrict__ start_idx, int32_t* __restrict__ out_idx) {
  extern __shared__ float sm[];
  __shared__ float rv[32];
  __shared__ int ri[32];
  __shared__ int s_far;
  int b = blockIdx.x, s = cloud_off[b], n = cloud_off[b + 1] - s;
  float* sx = sm; float* sy = sx + n; float* sz = sy + n; float* sd = sz + n;
  if (row_stride == 3) {
    // packed xyz: the cloud is one contiguous run of 3 n floats -> coalesced loads, de-interleaved into the SoA staging
    const float* src = xyz + (int64_t)s * 3;
    for (int f = threadIdx.x; f < 3 * n; f += blockDim.x) {
      const float v = src[f];
      const int pi = f / 3, c = f - 3 * pi;
      (c == 0 ? sx : c == 1 ? sy : sz)[pi] = v;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) sd[i] = 1e10f;
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float* p = xyz + (int64_t)(s + i) * row_stride;       // xyz columns of a wider feature row (pc_fts): 12 of every row_stride*4 bytes
      sx[i] = p[0]; sy[i] = p[1]; sz[i] = p[2];
      sd[i] = 1e10f;
    }
  }
  if (threadIdx.x == 0) s_far = start_idx[b];
  __syncthreads();
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int it = 0; it < npoint; ++it) {
    int far = s_far;
    if (threadIdx.x == 0) out_idx[(int64_t)b * npoint + it] = s + far;
    float cx = sx[far], cy = sy[far], cz = sz[far];
    float best = -1.f; int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      float dx = __fsub_rn(sx[i], cx), dy = __fsub_rn(sy[i], cy), dz = __fsub_rn(sz[i], cz);
      float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      float cur = sd[i];
      if (d < cur) { cur = d; sd[i] = d; }
      if (cur > best) { best = cur; bi = i; }  // ascending i within a thread: first max kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    __syncthreads();  // everyone has read s_far
    if (lane == 0) { rv[w] = best; ri[w] = bi; }
    __syncthreads();
    if (w == 0) {
      best = lane < nw ? rv[lane] : -1.f;
      bi = lane < nw ? ri[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) s_far = bi;
    }
    __syncthreads();
  }
}

LOTUS_API int lotus_fps(const float* xyz, int64_t row_stride, const int32_t* cloud_off, int B, int max_cloud,
                        int npoint, const int32_t* start_idx, int32_t* out_idx, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  size_t smem = (size_t)max_cloud * 16;
  LOTUS_CHECK_ARG(smem <= 220 * 1024, "lotus_fps: cloud of %d points exceeds the shared-memory staging (max 14080)", max_cloud);
  LOTUS_CUDA_CALL(cudaFuncSetAttribute(k_fps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_fps<<<B, 512, smem, st>>>(xyz, row_stride, cloud_off, npoint, start_idx, out_idx);
  LOTUS_CUDA_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// brute-force kNN / ball query inside each cloud: reference points tiled through shared memory
// ------------------------------------------------------------------------------------------------
#define KNN_TILE 1024
// coalesced staging of one reference tile: the tile is a contiguous run of 3 tn floats; stored as float4 (x, y, z, -) so the scan
// loop reads one candidate with ONE broadcast 16-byte shared load
__device__ __forceinline__ void knn_stage_tile(const float* __restrict__ src, int tn, float4* tile) {
  float* flat = (float*)tile;
  for (int f = threadIdx.x; f < 3 * tn; f += blockDim.x) {
    const int pi = f / 3, c = f - 3 * pi;
    flat[pi * 4 + c] = src[f];
  }
}
template <int KMAX>
__global__ void __launch_bounds__(128) k_knn(const float* __restrict__ q, const int32_t* __restrict__ q_off, const float* __restrict__ r,
                      const int32_t* __restrict__ r_off, int k, int32_t* __restrict__ idx_out,
                      float* __restrict__ dist_out) {
  __shared__ float4 tile[KNN_TILE];
  int b = blockIdx.y;
  int qs = q_off[b], qn = q_off[b + 1] - qs, rs = r_off[b], rn = r_off[b + 1] - rs;
  int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x * blockDim.x >= qn) return;
  bool active = qi < qn;
  float qx = 0, qy = 0, qz = 0;
  if (active) { qx = q[(int64_t)(qs + qi) * 3]; qy = q[(int64_t)(qs + qi) * 3 + 1]; qz = q[(int64_t)(qs + qi) * 3 + 2]; }
  float bd[KMAX]; int bi[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { bd[j] = INFINITY; bi[j] = -1; }
  float thr = INFINITY;                        // current k-th best distance, kept in a register: the scan loop is compare-only
  for (int t0 = 0; t0 < rn; t0 += KNN_TILE) {
    int tn = min(KNN_TILE, rn - t0);
    __syncthreads();
    knn_stage_tile(r + (int64_t)(rs + t0) * 3, tn, tile);
    __syncthreads();
    if (!active) continue;
#pragma unroll 4
    for (int i = 0; i < tn; ++i) {
      const float4 p = tile[i];
      float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
      float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d < thr) {
        // bubble the candidate through the sorted list with STATIC register indexing (a data-dependent shift loop put the list in
        // local memory, and every insertion of any lane stalled its whole warp on it): the candidate takes the first slot whose entry
        // is strictly larger (equal distances keep the earlier, lower index in front), everything behind shifts, the largest falls out
        float cd = d; int ci = rs + t0 + i;
        bool sw = false;                                       // sticky: past the insertion slot every entry just shifts right
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          sw = sw || cd < bd[j];
          const float td = bd[j]; const int ti = bi[j];
          bd[j] = sw ? cd : td; bi[j] = sw ? ci : ti;
          cd = sw ? td : cd; ci = sw ? ti : ci;
        }
        thr = bd[KMAX - 1];
      }
    }
  }
  if (active) {
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
      if (j < k) {
        idx_out[(int64_t)(qs + qi) * k + j] = bi[j];
        if (dist_out) dist_out[(int64_t)(qs + qi) * k + j] = bd[j];
      }
  }
}

LOTUS_API int lotus_knn(const float* query, const int32_t* q_off, const float* ref, const int32_t* r_off, int B,
                        int max_q_cloud, int k, int32_t* idx_out, float* dist_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  LOTUS_CHECK_ARG(k >= 1 && k <= 64, "lotus_knn: k=%d (1..64)", k);
  dim3 grid(ceil_div(max_q_cloud, 128), B);
  if (k <= 16) k_knn<16><<<grid, 128, 0, st>>>(query, q_off, ref, r_off, k, idx_out, dist_out);
  else if (k <= 32) k_knn<32><<<grid, 128, 0, st>>>(query, q_off, ref, r_off, k, idx_out, dist_out);
  else k_knn<64><<<grid, 128, 0, st>>>(query, q_off, ref, r_off, k, idx_out, dist_out);
  LOTUS_CUDA_OK();
  return 0;
}

__global__ void k_ball(const float* __restrict__ q, const int32_t* __restrict__ q_off, const float* __restrict__ r,
                       const int32_t* __restrict__ r_off, float radius2, int nsample, int32_t* __restrict__ idx_out) {
  __shared__ float4 tile[KNN_TILE];
  int b = blockIdx.y;
  int qs = q_off[b], qn = q_off[b + 1] - qs, rs = r_off[b], rn = r_off[b + 1] - rs;
  int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x * blockDim.x >= qn) return;
  bool active = qi < qn;
  float qx = 0, qy = 0, qz = 0;
  if (active) { qx = q[(int64_t)(qs + qi) * 3]; qy = q[(int64_t)(qs + qi) * 3 + 1]; qz = q[(int64_t)(qs + qi) * 3 + 2]; }
  int cnt = 0, first = -1;
  int32_t* out = idx_out + (int64_t)(qs + qi) * nsample;
  for (int t0 = 0; t0 < rn; t0 += KNN_TILE) {
    int tn = min(KNN_TILE, rn - t0);
    __syncthreads();
    knn_stage_tile(r + (int64_t)(rs + t0) * 3, tn, tile);
    __syncthreads();
    if (!active || cnt >= nsample) continue;
    for (int i = 0; i < tn && cnt < nsample; ++i) {
      const float4 p = tile[i];
      float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
      float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d < radius2) {
        if (cnt == 0) first = rs + t0 + i;
        out[cnt++] = rs + t0 + i;
      }
    }
  }
  if (active)
    for (int j = cnt; j < nsample; ++j) out[j] = first;
}

LOTUS_API int lotus_ball_query(const float* query, const int32_t* q_off, const float* ref, const int32_t* r_off, int B,
                               int max_q_cloud, float radius, int nsample, int32_t* idx_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  LOTUS_CHECK_ARG(nsample >= 1, "lotus_ball_query: nsample=%d", nsample);
  dim3 grid(ceil_div(max_q_cloud, 128), B);
  k_ball<<<grid, 128, 0, st>>>(query, q_off, ref, r_off, radius * radius, nsample, idx_out);
  LOTUS_CUDA_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// relative-position bias (RPE.forward, model.py:317-326) materialised: bias[p,h,i,j]
// ------------------------------------------------------------------------------------------------
__global__ void k_rpe_bias(const float* __restrict__ table, const int32_t* __restrict__ grid,
                           const int32_t* __restrict__ gather_idx, int K, int H, int pos_bnd,
                           float* __restrict__ bias) {
  int p = blockIdx.x;
  int num = 2 * pos_bnd + 1;
  for (int t = threadIdx.x; t < K * K; t += blockDim.x) {
    int i = t / K, j = t % K;
    int gi = gather_idx[p * K + i], gj = gather_idx[p * K + j];
    int idx[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      int d = grid[gi * 3 + a] - grid[gj * 3 + a];
      d = max(-pos_bnd, min(pos_bnd, d));
      idx[a] = d + pos_bnd + a * num;
    }
    for (int h = 0; h < H; ++h)
      bias[(((int64_t)p * H + h) * K + i) * K + j] = table[idx[0] * H + h] + table[idx[1] * H + h] + table[idx[2] * H + h];
  }
}
LOTUS_API int lotus_rpe_bias(const float* table, const int32_t* grid, const int32_t* gather_idx, int n_patches, int K,
                             int H, int pos_bnd, float* bias, void* stream_) {
  k_rpe_bias<<<n_patches, 256, 0, (cudaStream_t)stream_>>>(table, grid, gather_idx, K, H, pos_bnd, bias);
  LOTUS_CUDA_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// compact (output voxel, input voxel) pair lists per kernel tap — the rulebook in spconv's sense.
// pairs_out[d*stride + j] = p, pairs_in[d*stride + j] = nbr[d][p] for the j-th voxel (ascending p) that has
// a neighbour at tap d; pair_cnt[d] = number of such voxels.  Used by the weight-gradient kernels so they
// only contract over existing pairs.
// ------------------------------------------------------------------------------------------------
#define PAIR_CHUNK 1024
__global__ void k_pair_count(const int32_t* __restrict__ nbr, int64_t nbr_stride, const int32_t* __restrict__ meta,
                             int nchunks, int32_t* __restrict__ chunk_cnt /*[K3][nchunks]*/) {
  __shared__ int red[32];
  int d = blockIdx.y, ch = blockIdx.x, n = meta[META_N];
  int p0 = ch * PAIR_CHUNK;
  int c = 0;
  for (int i = threadIdx.x; i < PAIR_CHUNK; i += blockDim.x) {
    int p = p0 + i;
    if (p < n && nbr[(int64_t)d * nbr_stride + p] >= 0) ++c;
  }
  c = warp_sum_i(c);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0;
    v = warp_sum_i(v);
    if (threadIdx.x == 0) chunk_cnt[d * nchunks + ch] = v;
  }
}
__global__ void k_pair_scan(int32_t* __restrict__ chunk_cnt, int nchunks, int32_t* __restrict__ pair_cnt) {
  int d = blockIdx.x;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int c = 0; c < nchunks; ++c) {
      int v = chunk_cnt[d * nchunks + c];
      chunk_cnt[d * nchunks + c] = acc;
      acc += v;
    }
    pair_cnt[d] = acc;
  }
}
__global__ void k_pair_fill(const int32_t* __restrict__ nbr, int64_t nbr_stride, const int32_t* __restrict__ meta,
                            int nchunks, const int32_t* __restrict__ chunk_off, int32_t* __restrict__ pairs_out,
                            int32_t* __restrict__ pairs_in) {
  __shared__ int wsum[32];
  __shared__ int carry;
  int d = blockIdx.y, ch = blockIdx.x, n = meta[META_N];
  int p0 = ch * PAIR_CHUNK;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) carry = chunk_off[d * nchunks + ch];
  __syncthreads();
  for (int i0 = 0; i0 < PAIR_CHUNK; i0 += blockDim.x) {
    int p = p0 + i0 + threadIdx.x;
    int q = (p < n) ? nbr[(int64_t)d * nbr_stride + p] : -1;
    int f = q >= 0 ? 1 : 0, v = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) wsum[w] = v;
    __syncthreads();
    if (w == 0) {
      int t = lane < nw ? wsum[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u;
      }
      wsum[lane] = t;
    }
    __syncthreads();
    int incl = v + (w > 0 ? wsum[w - 1] : 0) + carry;
    if (f) {
      pairs_out[(int64_t)d * nbr_stride + incl - 1] = p;
      pairs_in[(int64_t)d * nbr_stride + incl - 1] = q;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = incl;
    __syncthreads();
  }
}
// scratch: K3 * ceil(n_upper / 1024) ints
LOTUS_API int lotus_pair_lists(const int32_t* nbr, int64_t nbr_stride, const int32_t* meta, int n_upper, int K3,
                               int32_t* pairs_out, int32_t* pairs_in, int32_t* pair_cnt, int32_t* scratch, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  int nchunks = ceil_div(n_upper, PAIR_CHUNK);
  k_pair_count<<<dim3(nchunks, K3), 256, 0, st>>>(nbr, nbr_stride, meta, nchunks, scratch);
  k_pair_scan<<<K3, 32, 0, st>>>(scratch, nchunks, pair_cnt);
  k_pair_fill<<<dim3(nchunks, K3), 256, 0, st>>>(nbr, nbr_stride, meta, nchunks, scratch, pairs_out, pairs_in);
  LOTUS_CUDA_OK();
  return 0;
}
