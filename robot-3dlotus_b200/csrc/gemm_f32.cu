// fp32 SIMT GEMM family — the bit-faithful "validation mode" of every dense contraction on the path
// (SURVEY §7 hard part "fp32 1e-4 parity vs bf16 speed").  One tiled kernel (64x64x16, 4x4 per thread)
// instantiated over operand loaders:
//   linear fwd / dgrad / wgrad ........ nn.Linear (model.py:386-387,572-574; model_ca.py:27-31)
//   submanifold conv fwd / dgrad / wgrad as an implicit GEMM over the d-major neighbour table
//                                        (SubMConv3d, model.py:616-622,845-852; spconv semantics restated)
// The tensor-core (tcgen05) path in gemm_tc.cu computes the same contractions in bf16.
#include "common.cuh"

#define BM 64
#define BN 64
#define BK 16

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_LEAKY = 2 };

struct Epi {
  float* out;           // [M, ldc]
  int64_t ldc;
  const float* bias;    // [N] or null
  const float* residual;  // [M, ldc] or null: out = residual + act(acc + bias)
  int act;
  int atomic;           // 1: atomicAdd into out (split-K), no bias/act
  float leaky;
  float drop_p;         // dropout on act(acc+bias) before the residual add (mask index = m*ldc+n)
  uint64_t seed;
  uint32_t drop_stream;
};

// ---- operand loaders: float operator()(row, k) with bounds already checked by the caller ----
struct ARowMajor {   // A[m, k] = p[m*ld + k]
  const float* p; int64_t ld;
  static constexpr bool kContig = true;
  __device__ __forceinline__ float operator()(int m, int k) const { return p[(int64_t)m * ld + k]; }
};
struct AColMajor {   // A[m, k] = p[k*ld + m]
  const float* p; int64_t ld;
  static constexpr bool kContig = false;
  __device__ __forceinline__ float operator()(int m, int k) const { return p[(int64_t)k * ld + m]; }
};
struct AGatherRows {  // A[m, k] = p[idx[m]*ld + k], idx<0 -> 0     (gathered linear input)
  const float* p; int64_t ld; const int32_t* idx;
  static constexpr bool kContig = true;
  __device__ __forceinline__ float operator()(int m, int k) const {
    int r = idx[m];
    return r < 0 ? 0.f : p[(int64_t)r * ld + k];
  }
};
struct AConv {       // A[m, kk=(d,c)] = x[nbr[dmap(d)*stride + m]*C + c]  (implicit-GEMM gather)
  const float* x; const int32_t* nbr; int64_t nbr_stride; int C; int K3; int mirror;
  static constexpr bool kContig = true;
  __device__ __forceinline__ float operator()(int m, int kk) const {
    int d = kk / C, c = kk - d * C;
    if (mirror) d = K3 - 1 - d;
    int r = nbr[(int64_t)d * nbr_stride + m];
    return r < 0 ? 0.f : x[(int64_t)r * C + c];
  }
};
struct BRowMajorT {  // B[k, n] = w[n*ld + k]   (weight [N,K], k contiguous)
  const float* p; int64_t ld;
  static constexpr bool kContig = true;
  __device__ __forceinline__ float operator()(int k, int n) const { return p[(int64_t)n * ld + k]; }
};
struct BRowMajor {   // B[k, n] = p[k*ld + n]
  const float* p; int64_t ld;
  static constexpr bool kContig = false;
  __device__ __forceinline__ float operator()(int k, int n) const { return p[(int64_t)k * ld + n]; }
};
struct BConvDgrad {  // B[kk=(d,co), n=ci] = w[(co*K3 + d)*Cin + ci]
  const float* w; int Cin; int Cout; int K3;
  static constexpr bool kContig = false;
  __device__ __forceinline__ float operator()(int kk, int n) const {
    int d = kk / Cout, co = kk - d * Cout;
    return w[((int64_t)co * K3 + d) * Cin + n];
  }
};
struct BGatherRows {  // B[k=p, n] = x[idx[p]*ld + n], idx<0 -> 0
  const float* p; int64_t ld; const int32_t* idx;
  static constexpr bool kContig = false;
  __device__ __forceinline__ float operator()(int k, int n) const {
    int r = idx[k];
    return r < 0 ? 0.f : p[(int64_t)r * ld + n];
  }
};

template <class LA, class LB>
__global__ void __launch_bounds__(256) k_gemm(int M, int N, int K, LA la, LB lb, Epi ep, int k_per_split) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  int tid = threadIdx.x;
  int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int kbeg = blockIdx.z * k_per_split, kend = min(K, kbeg + k_per_split);
  int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 4x4
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- stage A tile (BM x BK) ----
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int mm, kk;
      if (LA::kContig) { kk = tid & 15; mm = (tid >> 4) + 16 * r; }
      else { mm = tid & 63; kk = (tid >> 6) + 4 * r; }
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < kend) ? la(gm, gk) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int nn, kk;
      if (LB::kContig) { kk = tid & 15; nn = (tid >> 4) + 16 * r; }
      else { nn = tid & 63; kk = (tid >> 6) + 4 * r; }
      int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < kend) ? lb(gk, gn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
      *(float4*)a = *(const float4*)&As[kk][ty * 4];
      *(float4*)b = *(const float4*)&Bs[kk][tx * 4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      int64_t o = (int64_t)gm * ep.ldc + gn;
      if (ep.atomic) {
        atomicAdd(&ep.out[o], v);
      } else {
        if (ep.bias) v += ep.bias[gn];
        if (ep.act == ACT_GELU) v = gelu_f(v);
        else if (ep.act == ACT_LEAKY) v = v > 0.f ? v : v * ep.leaky;
        if (ep.drop_p > 0.f) {
          uint32_t thr = drop_thr(ep.drop_p);
          v = drop_keep_e(ep.seed, ep.drop_stream, (uint64_t)o, thr) ? v * drop_scale(thr) : 0.f;
        }
        if (ep.residual) v += ep.residual[o];
        ep.out[o] = v;
      }
    }
  }
}

template <class LA, class LB>
static int launch_gemm(int M, int N, int K, LA la, LB lb, Epi ep, int splits, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  int kps = K;
  if (splits > 1) {
    kps = ((K + splits - 1) / splits + BK - 1) / BK * BK;
    splits = (K + kps - 1) / kps;
  } else {
    splits = 1;
  }
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM), splits);
  k_gemm<LA, LB><<<grid, 256, 0, st>>>(M, N, K, la, lb, ep, kps);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { lotus_set_error("gemm_f32 launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

static int pick_splits(int M_out, int N_out, int K_red) {
  int tiles = ceil_div(M_out, BM) * ceil_div(N_out, BN);
  int want = (kNumSM * 4 + tiles - 1) / tiles;
  int maxs = max(1, K_red / 256);
  return max(1, min(want, maxs));
}

// ---- nn.Linear ----
// y[M,N] = act(x[M,K] @ w[N,K]^T + bias) (+ residual).  gather (optional): x row m is x[gather[m]] (row <0 -> zeros)
LOTUS_API int lotus_linear_fwd_f32(const float* x, const float* w, const float* bias, const float* residual, float* y,
                                   int M, int N, int K, int act, const int32_t* gather, float drop_p, uint64_t seed,
                                   uint32_t drop_stream, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  Epi ep{y, N, bias, residual, act, 0, 0.02f, drop_p, seed, drop_stream};
  if (gather) return launch_gemm(M, N, K, AGatherRows{x, K, gather}, BRowMajorT{w, K}, ep, 1, st);
  return launch_gemm(M, N, K, ARowMajor{x, K}, BRowMajorT{w, K}, ep, 1, st);
}
// dx[M,K] = dy[M,N] @ w[N,K]   (+ residual: dx = residual + ...)
LOTUS_API int lotus_linear_dgrad_f32(const float* dy, const float* w, const float* residual, float* dx, int M, int N,
                                     int K, void* stream_) {
  Epi ep{dx, K, nullptr, residual, ACT_NONE, 0, 0.f, 0.f, 0, 0};
  return launch_gemm(M, K, N, ARowMajor{dy, N}, BRowMajor{w, K}, ep, 1, (cudaStream_t)stream_);
}
// dw[N,K] += dy[M,N]^T @ x[M,K]  (dw must be zero-initialised or hold the running gradient); gather as in fwd
LOTUS_API int lotus_linear_wgrad_f32(const float* dy, const float* x, float* dw, int M, int N, int K,
                                     const int32_t* gather, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  Epi ep{dw, K, nullptr, nullptr, ACT_NONE, 1, 0.f, 0.f, 0, 0};
  int splits = pick_splits(N, K, M);
  if (gather) return launch_gemm(N, K, M, AColMajor{dy, N}, BGatherRows{x, K, gather}, ep, splits, st);
  return launch_gemm(N, K, M, AColMajor{dy, N}, BRowMajor{x, K}, ep, splits, st);
}

// ---- submanifold conv as implicit GEMM over the d-major neighbour table nbr[K3][nbr_stride] ----
// y[n, Cout] = sum_{d,ci} x[nbr[d][n], ci] * w[co, d, ci] (+ bias)
LOTUS_API int lotus_subm_conv_fwd_f32(const float* x, const float* w, const float* bias, float* y, const int32_t* nbr,
                                      int64_t nbr_stride, int n, int Cin, int Cout, int K3, void* stream_) {
  Epi ep{y, Cout, bias, nullptr, ACT_NONE, 0, 0.f, 0.f, 0, 0};
  return launch_gemm(n, Cout, K3 * Cin, AConv{x, nbr, nbr_stride, Cin, K3, 0}, BRowMajorT{w, (int64_t)K3 * Cin}, ep, 1,
                     (cudaStream_t)stream_);
}
// dx[q, ci] = [residual[q, ci] +] sum_{d,co} dy[nbr[K3-1-d][q], co] * w[co, d, ci]
LOTUS_API int lotus_subm_conv_dgrad_f32(const float* dy, const float* w, const float* residual, float* dx, const int32_t* nbr,
                                        int64_t nbr_stride, int n, int Cin, int Cout, int K3, void* stream_) {
  Epi ep{dx, Cin, nullptr, residual, ACT_NONE, 0, 0.f, 0.f, 0, 0};
  return launch_gemm(n, Cin, K3 * Cout, AConv{dy, nbr, nbr_stride, Cout, K3, 1}, BConvDgrad{w, Cin, Cout, K3}, ep, 1,
                     (cudaStream_t)stream_);
}
// narrow-input weight gradient (the stem: Cin = 7 / 4 / 68-ish is handled by the generic path, Cin*Cout <= 1024 here):
// one CTA per (point chunk, tap): compact the chunk's existing pairs in shared memory, then thread (co, ci)
// accumulates sum_j dy[p_j, co] * x[q_j, ci] and issues one atomicAdd.
#define STEM_CHUNK 2048
__global__ void k_conv_wgrad_narrow(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                                    const int32_t* __restrict__ nbr, int64_t nbr_stride, int n, int Cin, int Cout, int K3) {
  __shared__ int lp[STEM_CHUNK], lq[STEM_CHUNK];
  __shared__ int cnt;
  int d = blockIdx.y, p0 = blockIdx.x * STEM_CHUNK;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < STEM_CHUNK; i += blockDim.x) {
    int p = p0 + i;
    if (p < n) {
      int q = nbr[(int64_t)d * nbr_stride + p];
      if (q >= 0) { int pos = atomicAdd(&cnt, 1); lp[pos] = p; lq[pos] = q; }
    }
  }
  __syncthreads();
  int t = threadIdx.x;
  if (t >= Cin * Cout) return;
  int co = t % Cout, ci = t / Cout, m = cnt;
  float acc = 0.f;
  for (int j = 0; j < m; ++j) acc = fmaf(dy[(int64_t)lp[j] * Cout + co], x[(int64_t)lq[j] * Cin + ci], acc);
  if (m > 0) atomicAdd(&dw[((int64_t)co * K3 + d) * Cin + ci], acc);
}

// dw[co, d, ci] += sum_p dy[p, co] * x[nbr[d][p], ci]
LOTUS_API int lotus_subm_conv_wgrad_f32(const float* dy, const float* x, float* dw, const int32_t* nbr,
                                        int64_t nbr_stride, int n, int Cin, int Cout, int K3, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (Cin * Cout <= 1024 && n > 0) {
    int threads = ((Cin * Cout + 31) / 32) * 32;
    k_conv_wgrad_narrow<<<dim3(ceil_div(n, STEM_CHUNK), K3), threads, 0, st>>>(dy, x, dw, nbr, nbr_stride, n, Cin, Cout, K3);
    LOTUS_CUDA_OK();
    return 0;
  }
  int splits = pick_splits(Cout, Cin, n);
  splits = max(1, splits / 4);
  for (int d = 0; d < K3; ++d) {
    Epi ep{dw + (int64_t)d * Cin, (int64_t)K3 * Cin, nullptr, nullptr, ACT_NONE, 1, 0.f, 0.f, 0, 0};
    int rc = launch_gemm(Cout, Cin, n, AColMajor{dy, Cout}, BGatherRows{x, Cin, nbr + (int64_t)d * nbr_stride}, ep,
                         splits, st);
    if (rc) return rc;
  }
  return 0;
}
