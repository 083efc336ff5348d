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

// ---- the stem: narrow input (Cin <= 8), Cout % 64 == 0, 5x5x5 taps.  fp32 SIMT kernels that exploit the table's sparsity ----
// As an implicit GEMM the stem is K = 125 * Cin of which ~65 % multiplies zero-filled rows (absent neighbours), and the scalar
// gather of the generic loader above costs more than the FMAs.  Here a warp owns 32 consecutive points (neighbouring rows of the
// table, so the tap look-up is one coalesced load + ballot), a lane owns two output channels (co = lane, lane + 32), and the
// (point, tap) pair is the unit of work: the neighbour's Cin values are warp-uniform loads, the 2 * Cin weights of the tap sit in
// registers, absent pairs cost a mask test.  The weights of a group of taps are staged through shared memory, transposed to
// [tap][ci][co] (row pitch 65: conflict-free both ways).  Same arithmetic order per output as the generic kernel up to the order of
// the tap sum (fp32, within the 1e-4 parity bound).
__host__ __device__ constexpr int stem_tg(int cin) { return 25; }   // taps per shared-memory group
// The neighbour rows of a tap travel global -> per-warp shared-memory slab ring [3][32][8] by cp.async (4 bytes x Cin per present
// lane), issued TWO taps ahead, and are read back as warp-uniform LDS.128 in the pair loop; the table rows run four taps ahead in
// registers.  Without the look-ahead every tap paid an L2 round trip in series (0.55 ms for the 96k-point stem).
__device__ __forceinline__ void stem_cp4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void stem_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void stem_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
// weights of `ntap` taps -> sW[(t * CIN + ci) * 65 + co] = w[((co0 + co) * K3 + d0 + t) * CIN + ci], all copies in flight at once
// (the caller drains with wait_group 0 + __syncthreads)
template <int CIN>
__device__ __forceinline__ void stem_stage_weights(const float* __restrict__ w, float* sW, int co0, int K3, int d0, int ntap) {
  const int per_co = ntap * CIN;
  for (int co = threadIdx.x >> 5; co < 64; co += blockDim.x >> 5) {          // a warp walks one output channel's contiguous run
    const float* src = w + ((int64_t)(co0 + co) * K3 + d0) * CIN;
    for (int k = threadIdx.x & 31; k < per_co; k += 32) stem_cp4(sW + k * 65 + co, src + k);
  }
  stem_commit();
}
template <int CIN>
__device__ __forceinline__ void stem_fetch_row(const float* __restrict__ x, int q, float* slab_row) {
  if (q >= 0) {
    const float* xr = x + (int64_t)q * CIN;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) stem_cp4(slab_row + ci, xr + ci);
  }
  stem_commit();
}
template <int CIN>
__device__ __forceinline__ void stem_get_row(const float* slab, int i, float* xv) {
  const float4 a = *(const float4*)(slab + i * 8);
  xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
  if (CIN > 4) { const float4 b = *(const float4*)(slab + i * 8 + 4); xv[4] = b.x; xv[5] = b.y; xv[6] = b.z; xv[7] = b.w; }
}
#define STEM_SLAB 768                    // floats per warp: [3][32][8]
template <int CIN>
__global__ void __launch_bounds__(256, 2) k_conv_fwd_narrow(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ y,
                                                            const int32_t* __restrict__ nbr, int64_t nbr_stride, int n, int Cout, int K3) {
  constexpr int TG = stem_tg(CIN);
  extern __shared__ __align__(16) float stem_smem[];
  float* sW = stem_smem;                                         // [TG * CIN][65]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* slab = stem_smem + ((TG * CIN * 65 + 3) & ~3) + warp * STEM_SLAB;
  const int p0 = (blockIdx.x * 8 + warp) * 32;
  const int co0 = blockIdx.y * 64;
  const bool tile_ok = p0 < n, pt_ok = p0 + lane < n;
  const int32_t* nb = nbr + p0 + lane;
  float acc[32][2];
  {
    const float b0 = bias ? bias[co0 + lane] : 0.f, b1 = bias ? bias[co0 + 32 + lane] : 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { acc[i][0] = b0; acc[i][1] = b1; }
  }
  const int32_t* nb_look = nb;                                   // table row of the next tap to request
  auto tap_row = [&](bool in_range) { const int v = (pt_ok && in_range) ? *nb_look : -1; nb_look += nbr_stride; return v; };
  int q0 = tap_row(0 < K3), q1 = tap_row(1 < K3), q2 = tap_row(2 < K3), q3 = tap_row(3 < K3);
  stem_fetch_row<CIN>(x, q0, slab + lane * 8);
  stem_fetch_row<CIN>(x, q1, slab + 256 + lane * 8);
  for (int d = 0; d < K3; ++d) {
    const int t = d % TG;
    if (t == 0) {
      __syncthreads();                                           // every warp is done with the previous group's weights
      stem_stage_weights<CIN>(w, sW, co0, K3, d, min(TG, K3 - d));
      asm volatile("cp.async.wait_group 0;" ::: "memory");      // (also drains the two row groups in flight: they are re-waited below for free)
      __syncthreads();
    }
    const int q4 = tap_row(d + 4 < K3);
    stem_wait1();                                                // this lane's rows of tap d have landed
    __syncwarp();                                                // ... every lane's; and every lane is done with tap d - 1
    stem_fetch_row<CIN>(x, q2, slab + ((d + 2) % 3) * 256 + lane * 8);
    const unsigned m = __ballot_sync(0xffffffffu, q0 >= 0);
    if (m != 0 && tile_ok) {
      const float* sl = slab + (d % 3) * 256;
      float w0[CIN], w1[CIN];
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) { w0[ci] = sW[(t * CIN + ci) * 65 + lane]; w1[ci] = sW[(t * CIN + ci) * 65 + 32 + lane]; }
      // static point index (the accumulators keep their register names); groups of four absent points cost one test
#pragma unroll
      for (int i4 = 0; i4 < 8; ++i4) {
        if ((m >> (4 * i4)) & 0xFu) {                            // warp-uniform
#pragma unroll
          for (int i = 4 * i4; i < 4 * i4 + 4; ++i) {
            if (m & (1u << i)) {
              float xv[8];
              stem_get_row<CIN>(sl, i, xv);
#pragma unroll
              for (int ci = 0; ci < CIN; ++ci) {
                acc[i][0] = fmaf(xv[ci], w0[ci], acc[i][0]);
                acc[i][1] = fmaf(xv[ci], w1[ci], acc[i][1]);
              }
            }
          }
        }
      }
    }
    q0 = q1; q1 = q2; q2 = q3; q3 = q4;
  }
  if (!tile_ok) return;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (p0 + i < n) {
      float* yr = y + (int64_t)(p0 + i) * Cout + co0;
      yr[lane] = acc[i][0];
      yr[32 + lane] = acc[i][1];
    }
  }
}

// dw[co, d, ci] += sum_p dy[p, co] * x[nbr[d][p], ci].  CTA = (point chunk, tap residue g: taps d = g, g + groups, ...): a warp keeps
// the dy rows of its 32 points in registers across the CTA's taps, forms the tap's partial in registers, and adds it to the CTA's
// shared-memory accumulator; one global atomic per (tap, ci, co) per CTA at the end.
#define STEM_WG_WARPS 16
template <int CIN>
__global__ void __launch_bounds__(STEM_WG_WARPS * 32, 1) k_conv_wgrad_narrow2(const float* __restrict__ dy, const float* __restrict__ x,
                                                                             float* __restrict__ dw, const int32_t* __restrict__ nbr,
                                                                             int64_t nbr_stride, int n, int Cout, int K3, int tiles_per_cta,
                                                                             int groups) {
  constexpr int TG = stem_tg(CIN);
  extern __shared__ __align__(16) float stem_smem[];
  float* sAcc = stem_smem;                                       // [TG * CIN][64]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* slab = stem_smem + TG * CIN * 64 + warp * STEM_SLAB;
  const int g = blockIdx.y, co0 = blockIdx.z * 64;
  const int ntap = (K3 - g + groups - 1) / groups;              // taps g, g + groups, ...
  for (int i = threadIdx.x; i < TG * CIN * 64; i += blockDim.x) sAcc[i] = 0.f;
  __syncthreads();
  const int tile0 = blockIdx.x * tiles_per_cta;
  for (int tl = warp; tl < tiles_per_cta; tl += STEM_WG_WARPS) {
    const int p0 = (tile0 + tl) * 32;
    if (p0 >= n) break;
    const bool pt_ok = p0 + lane < n;
    const int32_t* nb = nbr + p0 + lane;
    // taps are walked out of phase by the warps (fewer collisions on sAcc): step tt visits tap (tt + warp) % ntap
    const int t_first = warp % ntap;
    const int64_t tap_step = (int64_t)groups * nbr_stride;
    const int32_t* nb_first = nb + (int64_t)g * nbr_stride;      // table row of the group's first tap
    const int32_t* nb_look = nb_first + t_first * tap_step;      // ... of the next tap to request (wraps without a division)
    int t_look = t_first;
    auto next_row = [&](bool in_range) {
      const int v = (pt_ok && in_range) ? *nb_look : -1;
      nb_look += tap_step;
      if (++t_look == ntap) { t_look = 0; nb_look = nb_first; }
      return v;
    };
    int q0 = next_row(0 < ntap), q1 = next_row(1 < ntap), q2 = next_row(2 < ntap), q3 = next_row(3 < ntap);
    __syncwarp();                                                // the previous tile's reads of the slab are done
    stem_fetch_row<CIN>(x, q0, slab + lane * 8);
    stem_fetch_row<CIN>(x, q1, slab + 256 + lane * 8);
    float dyv[32][2];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const bool ok = p0 + i < n;
      const float* dr = dy + (int64_t)(p0 + i) * Cout + co0;
      const float2 dv = ok ? *(const float2*)(dr + 2 * lane) : make_float2(0.f, 0.f);   // this lane's channels: 2 lane, 2 lane + 1
      dyv[i][0] = dv.x;
      dyv[i][1] = dv.y;
    }
    int t = t_first;
    for (int tt = 0; tt < ntap; ++tt) {
      const int q4 = next_row(tt + 4 < ntap);
      stem_wait1();
      __syncwarp();
      stem_fetch_row<CIN>(x, q2, slab + ((tt + 2) % 3) * 256 + lane * 8);
      const unsigned m = __ballot_sync(0xffffffffu, q0 >= 0);
      if (m != 0) {
        const float* sl = slab + (tt % 3) * 256;
        float a0[CIN], a1[CIN];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) { a0[ci] = 0.f; a1[ci] = 0.f; }
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          if ((m >> (4 * i4)) & 0xFu) {
#pragma unroll
            for (int i = 4 * i4; i < 4 * i4 + 4; ++i) {
              if (m & (1u << i)) {
                float xv[8];
                stem_get_row<CIN>(sl, i, xv);
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) {
                  a0[ci] = fmaf(xv[ci], dyv[i][0], a0[ci]);
                  a1[ci] = fmaf(xv[ci], dyv[i][1], a1[ci]);
                }
              }
            }
          }
        }
        // shared-memory fp32 atomicAdd is a CAS loop on this hardware: one 64-bit CAS covers the lane's channel pair
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          unsigned long long* addr = (unsigned long long*)&sAcc[(t * CIN + ci) * 64 + 2 * lane];
          unsigned long long old = *addr, assumed;
          do {
            assumed = old;
            const float lo = __uint_as_float((unsigned)assumed) + a0[ci], hi = __uint_as_float((unsigned)(assumed >> 32)) + a1[ci];
            old = atomicCAS(addr, assumed, ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo));
          } while (old != assumed);
        }
      }
      q0 = q1; q1 = q2; q2 = q3; q3 = q4;
      if (++t == ntap) t = 0;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");      // the look-ahead groups past the last tap are empty, but drain them
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ntap * CIN * 64; i += blockDim.x) {
    const int co = i / (ntap * CIN), k = i - co * (ntap * CIN);  // consecutive threads: consecutive (tap, ci) of one output channel
    const int t = k / CIN, ci = k - t * CIN;
    const float v = sAcc[(t * CIN + ci) * 64 + co];
    if (v != 0.f) atomicAdd(&dw[((int64_t)(co0 + co) * K3 + g + t * groups) * CIN + ci], v);
  }
}

template <int CIN>
static int stem_fwd_launch(const float* x, const float* w, const float* bias, float* y, const int32_t* nbr, int64_t nbr_stride, int n,
                           int Cout, int K3, cudaStream_t st) {
  const size_t smem = (size_t)(((stem_tg(CIN) * CIN * 65 + 3) & ~3) + 8 * STEM_SLAB) * 4;
  static bool set = false;
  if (!set) { LOTUS_CUDA_CALL(cudaFuncSetAttribute(k_conv_fwd_narrow<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = true; }
  k_conv_fwd_narrow<CIN><<<dim3(ceil_div(n, 256), Cout / 64), 256, smem, st>>>(x, w, bias, y, nbr, nbr_stride, n, Cout, K3);
  LOTUS_CUDA_OK();
  return 0;
}
template <int CIN>
static int stem_wgrad_launch(const float* dy, const float* x, float* dw, const int32_t* nbr, int64_t nbr_stride, int n, int Cout, int K3,
                             cudaStream_t st) {
  const int groups = ceil_div(K3, stem_tg(CIN));                 // 5 for 125 taps: 25 taps per CTA
  const int tiles = ceil_div(n, 32);
  int chunks = max(1, min(ceil_div(tiles, STEM_WG_WARPS), kNumSM / (groups * (Cout / 64))));
  const int tiles_per_cta = ceil_div(tiles, chunks);
  chunks = ceil_div(tiles, tiles_per_cta);
  const size_t smem = (size_t)(stem_tg(CIN) * CIN * 64 + STEM_WG_WARPS * STEM_SLAB) * 4;
  static bool set = false;
  if (!set) { LOTUS_CUDA_CALL(cudaFuncSetAttribute(k_conv_wgrad_narrow2<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = true; }
  k_conv_wgrad_narrow2<CIN><<<dim3(chunks, groups, Cout / 64), STEM_WG_WARPS * 32, smem, st>>>(dy, x, dw, nbr, nbr_stride, n, Cout, K3,
                                                                                                 tiles_per_cta, groups);
  LOTUS_CUDA_OK();
  return 0;
}
static inline bool stem_shape(int Cin, int Cout, int K3) { return Cin >= 1 && Cin <= 8 && Cout % 64 == 0 && K3 >= 1; }
#define STEM_DISPATCH(FN, ...)                                                                         \
  switch (Cin) {                                                                                       \
    case 1: return FN<1>(__VA_ARGS__); case 2: return FN<2>(__VA_ARGS__); case 3: return FN<3>(__VA_ARGS__); \
    case 4: return FN<4>(__VA_ARGS__); case 5: return FN<5>(__VA_ARGS__); case 6: return FN<6>(__VA_ARGS__); \
    case 7: return FN<7>(__VA_ARGS__); default: return FN<8>(__VA_ARGS__);                             \
  }

// ---- submanifold conv as implicit GEMM over the d-major neighbour table nbr[K3][nbr_stride] ----
// y[n, Cout] = sum_{d,ci} x[nbr[d][n], ci] * w[co, d, ci] (+ bias)
LOTUS_API int lotus_subm_conv_fwd_f32(const float* x, const float* w, const float* bias, float* y, const int32_t* nbr,
                                      int64_t nbr_stride, int n, int Cin, int Cout, int K3, void* stream_) {
  if (n <= 0) return 0;
  if (stem_shape(Cin, Cout, K3)) { STEM_DISPATCH(stem_fwd_launch, x, w, bias, y, nbr, nbr_stride, n, Cout, K3, (cudaStream_t)stream_) }
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
  if (n <= 0) return 0;
  if (stem_shape(Cin, Cout, K3)) { STEM_DISPATCH(stem_wgrad_launch, dy, x, dw, nbr, nbr_stride, n, Cout, K3, st) }
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
