// Patch self-attention on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators), bf16 activations.
// Same contract as the SIMT kernels in attention.cu (SerializedAttention.forward, model.py:468-557, flash branch):
// gather by `order[pad]`, scatter by `unpad[inverse]`, alias tail slots, fp32 softmax, counter-based dropout mask,
// lse = max + log(sum) of the scaled logits.  q and k arrive per-head LayerNorm'ed.
//
// One CTA = one (patch, head): a PTv3 patch is at most K = 128 tokens, i.e. exactly one UMMA M=128 tile.
//   S[128 x 128]  = Q[128 x Dh] . K[128 x Dh]^T        one tcgen05.mma chain (Dh/16 steps), accumulator in TMEM
//   softmax       : thread r owns query row r = TMEM lane r (tcgen05.ld 32x32b), P written to smem as bf16
//   O[128 x Dh]   = P[128 x 128] . V[128 x Dh]          second chain (8 steps), accumulator reuses the TMEM columns
// All operand tiles are K-major SWIZZLE_128B (row r at r*128 B, 16-byte chunk c at (c ^ (r & 7)) * 16), written by the
// threads themselves (the rows are gathered, so TMA tiles do not apply); V is stored transposed for the second chain.
#include "tc_common.cuh"
#include <stdlib.h>

namespace tca {
using namespace tc;

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *(uint32_t*)&v;
}

// instruction descriptor: D fp32, A/B bf16, both K-major unless flagged, N and M extents
__device__ __forceinline__ uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// DHP: head dim padded to a multiple of 16 (K extent of S, N extent of O); Dh <= DHP, Dh % 8 == 0
template <int DHP>
struct FwdSmem {
  static constexpr int Q_OFF = 0;                       // [128 rows][128 B]  (first DHP*2 bytes of each row used)
  static constexpr int K_OFF = 16384;
  static constexpr int VT_OFF = 32768;                  // 2 key-chunks x [DHP rows][128 B]
  static constexpr int P_OFF = VT_OFF + 2 * DHP * 128;  // 2 key-chunks x [128 rows][128 B]
  static constexpr int BAR_OFF = P_OFF + 32768;
  static constexpr int TOTAL = BAR_OFF + 64 + 1024;     // + alignment slack
};

template <int DHP>
__global__ void __launch_bounds__(128) k_patch_attn_fwd_tc(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v, int64_t ldq,
    int64_t ldk, int64_t ldv, const int32_t* __restrict__ gather_idx, const int32_t* __restrict__ slot_of_point,
    const int32_t* __restrict__ cu, int H, int Dh, float scale, __nv_bfloat16* __restrict__ out, int64_t ldo,
    float* __restrict__ lse, float drop_p, uint64_t seed, uint32_t layer_id) {
  using L = FwdSmem<DHP>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem + L::Q_OFF;
  uint8_t* sK = smem + L::K_OFF;
  uint8_t* sVt = smem + L::VT_OFF;
  uint8_t* sP = smem + L::P_OFF;
  uint64_t* bars = (uint64_t*)(smem + L::BAR_OFF);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2);

  const int p = blockIdx.x, h = blockIdx.y;
  const int s0 = cu[p], len = cu[p + 1] - s0;
  if (len <= 0) return;                                  // uniform per CTA
  const int r = threadIdx.x, warp = r >> 5;
  if (r == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);

  // ---- stage Q, K (rows) and V (transposed); rows >= len and head-dim padding are zero
  const bool live = r < len;
  const int pt = live ? gather_idx[s0 + r] : 0;
  const int nchunk = Dh >> 3;
  {
    const uint4* gq = (const uint4*)(q + (int64_t)pt * ldq + h * Dh);
    const uint4* gk = (const uint4*)(k + (int64_t)pt * ldk + h * Dh);
    const uint4* gv = (const uint4*)(v + (int64_t)pt * ldv + h * Dh);
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int c = 0; c < DHP / 8; ++c) {
      bool ok = live && c < nchunk;
      uint4 a = ok ? gq[c] : z, b = ok ? gk[c] : z, w = ok ? gv[c] : z;
      int sw = ((c ^ (r & 7)) << 4);
      *(uint4*)(sQ + r * 128 + sw) = a;
      *(uint4*)(sK + r * 128 + sw) = b;
      const __nv_bfloat16* wv = (const __nv_bfloat16*)&w;
      // V^T: element (d, key r) -> chunk-of-64-keys (r >> 6), row d, 16-B chunk ((r & 63) >> 3) ^ (d & 7), lane r & 7
      uint8_t* base = sVt + (r >> 6) * (DHP * 128) + (r & 7) * 2;
      int kc = (r & 63) >> 3;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int d = c * 8 + e;
        *(__nv_bfloat16*)(base + d * 128 + ((kc ^ (d & 7)) << 4)) = wv[e];
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- S = Q K^T
  if (r == 0) {
    const uint32_t idesc = make_idesc(128, 128, false, false);
    const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
#pragma unroll
    for (int kk = 0; kk < DHP / 16; ++kk)
      umma_bf16(tmem_base, make_desc(qa + kk * 32, 16, 1024), make_desc(ka + kk * 32, 16, 1024), idesc, kk > 0 ? 1u : 0u);
    umma_commit(&bars[0]);
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after();

  // ---- softmax of row r (TMEM lane r), dropout, P -> smem (bf16, unnormalised)
  float sv[128];
  {
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(lane_addr + c * 32, (uint32_t*)&sv[c * 32]);
    tmem_ld_wait();
  }
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < 128; ++j) {
    float t = j < len ? sv[j] * scale : -INFINITY;
    sv[j] = t;
    m = fmaxf(m, t);
  }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < 128; ++j) {
    float e = __expf(sv[j] - m);
    sv[j] = e;
    l += e;
  }
  if (drop_p > 0.f) {
    const float ks = 1.f / (1.f - drop_p);
    const uint32_t sid = layer_id * 0x9E37u + (uint32_t)(p * H + h);
#pragma unroll
    for (int j = 0; j < 128; ++j)
      sv[j] = (j < len && dropout_keep(seed, sid, (uint32_t)(r * len + j), drop_p)) ? sv[j] * ks : 0.f;
  }
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    uint4 w;
    w.x = pack_bf16x2(sv[c * 8 + 0], sv[c * 8 + 1]);
    w.y = pack_bf16x2(sv[c * 8 + 2], sv[c * 8 + 3]);
    w.z = pack_bf16x2(sv[c * 8 + 4], sv[c * 8 + 5]);
    w.w = pack_bf16x2(sv[c * 8 + 6], sv[c * 8 + 7]);
    *(uint4*)(sP + (c >> 3) * 16384 + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = w;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();            // every lane has drained S from TMEM (wait::ld above) and P is visible to the async proxy
  tc_fence_after();

  // ---- O = P V  (accumulator reuses TMEM columns [0, DHP))
  if (r == 0) {
    const uint32_t idesc = make_idesc(128, DHP, false, false);
    const uint32_t pa = smem_u32(sP), va = smem_u32(sVt);
#pragma unroll
    for (int kk = 0; kk < 8; ++kk)
      umma_bf16(tmem_base, make_desc(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                make_desc(va + (kk >> 2) * (DHP * 128) + (kk & 3) * 32, 16, 1024), idesc, kk > 0 ? 1u : 0u);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  {
    constexpr int NLD = (DHP + 31) / 32;
    float o[NLD * 32];
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < NLD; ++c) tmem_ld_32x32b_x32(lane_addr + c * 32, (uint32_t*)&o[c * 32]);
    tmem_ld_wait();
    if (live) {
      lse[(int64_t)(s0 + r) * H + h] = m + __logf(l);
      if (slot_of_point[pt] == s0 + r) {   // alias slots of a padded tail are queries whose output is discarded
        const float inv = 1.f / l;
        uint4* go = (uint4*)(out + (int64_t)pt * ldo + h * Dh);
#pragma unroll
        for (int c = 0; c < DHP / 8; ++c) {
          if (c < nchunk) {
            uint4 w;
            w.x = pack_bf16x2(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv);
            w.y = pack_bf16x2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
            w.z = pack_bf16x2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv);
            w.w = pack_bf16x2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
            go[c] = w;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

template <int DHP>
static int launch_fwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, const int32_t* gather_idx,
                      const int32_t* slot_of_point, const int32_t* cu, int n_patches, int H, int Dh, float scale, void* out,
                      int64_t ldo, float* lse, float drop_p, uint64_t seed, uint32_t layer_id, cudaStream_t st) {
  constexpr int SMEM = FwdSmem<DHP>::TOTAL;
  static bool attr_set = false;
  if (!attr_set) {
    LOTUS_CUDA_CALL(cudaFuncSetAttribute(k_patch_attn_fwd_tc<DHP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  k_patch_attn_fwd_tc<DHP><<<dim3(n_patches, H), 128, SMEM, st>>>(
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ldq, ldk, ldv, gather_idx, slot_of_point, cu, H, Dh,
      scale, (__nv_bfloat16*)out, ldo, lse, drop_p, seed, layer_id);
  LOTUS_CUDA_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------ backward
// One CTA (256 threads) = one (patch, head).  Everything is computed in the key-major (transposed) orientation so
// that the probability / score-gradient tiles land in shared memory directly as K-major A operands:
//   S^T [key x q]  = K Q^T          dP^T [key x q] = V dO^T                      (phase A, TMEM columns 0-127 / 128-255)
//   P^T = exp(S^T * scale - lse[q]) (dropout mask),  dS^T = P^T * (dP^T - delta[q]) * scale   (thread (key, q-half))
//   dV [key x Dh] = P^T dO          dK [key x Dh] = dS^T Q          dQ [q x Dh] = dS K      (phase B)
// dQ reads the dS^T tile as an MN-major A operand (same bytes, M = q contiguous).  dq / dk / dv are accumulated with
// vector atomics in fp32: a point that sits in an alias tail slot of the cloud's last patch contributes twice.
template <int DHP>
struct BwdSmem {
  static constexpr int Q_OFF = 0, K_OFF = 16384, V_OFF = 32768, DO_OFF = 49152;     // row tiles [128][128 B]
  static constexpr int TT = 2 * DHP * 128;                                          // transposed tile: 2 chunks x [DHP][128 B]
  static constexpr int QT_OFF = 65536, KT_OFF = QT_OFF + TT, DOT_OFF = KT_OFF + TT;
  static constexpr int PT_OFF = ((DOT_OFF + TT + 1023) / 1024) * 1024;              // 2 q-chunks x [128 keys][128 B]
  static constexpr int DST_OFF = PT_OFF + 32768;
  static constexpr int LSE_OFF = DST_OFF + 32768;                                   // float[128] lse, float[128] delta
  static constexpr int BAR_OFF = LSE_OFF + 1024;
  static constexpr int TOTAL = BAR_OFF + 64 + 1024;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// store one gathered row (DHP/8 chunks) into a K-major row tile and its transposed tile (either may be null)
template <int DHP>
__device__ __forceinline__ void stage_row(const uint4* __restrict__ src, bool live, int nchunk, int r, uint8_t* row_tile,
                                          uint8_t* t_tile, uint4* keep) {
  const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int c = 0; c < DHP / 8; ++c) {
    uint4 a = (live && c < nchunk) ? src[c] : z;
    if (keep) keep[c] = a;
    if (row_tile) *(uint4*)(row_tile + r * 128 + ((c ^ (r & 7)) << 4)) = a;
    if (t_tile) {
      const __nv_bfloat16* av = (const __nv_bfloat16*)&a;
      uint8_t* base = t_tile + (r >> 6) * (DHP * 128) + (r & 7) * 2;
      int kc = (r & 63) >> 3;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int d = c * 8 + e;
        *(__nv_bfloat16*)(base + d * 128 + ((kc ^ (d & 7)) << 4)) = av[e];
      }
    }
  }
}

template <int DHP>
__global__ void __launch_bounds__(256) k_patch_attn_bwd_tc(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v, int64_t ldq,
    int64_t ldk, int64_t ldv, const int32_t* __restrict__ gather_idx, const int32_t* __restrict__ slot_of_point,
    const int32_t* __restrict__ cu, int H, int Dh, float scale, const __nv_bfloat16* __restrict__ out,
    const __nv_bfloat16* __restrict__ dout, int64_t ldo, const float* __restrict__ lse, float* __restrict__ dq,
    float* __restrict__ dk, float* __restrict__ dv, int64_t lddq, int64_t lddk, int64_t lddv, float drop_p, uint64_t seed,
    uint32_t layer_id) {
  using L = BwdSmem<DHP>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *sQ = smem + L::Q_OFF, *sK = smem + L::K_OFF, *sV = smem + L::V_OFF, *sdO = smem + L::DO_OFF;
  uint8_t *sQt = smem + L::QT_OFF, *sKt = smem + L::KT_OFF, *sdOt = smem + L::DOT_OFF;
  uint8_t *sPt = smem + L::PT_OFF, *sdSt = smem + L::DST_OFF;
  float* sLse = (float*)(smem + L::LSE_OFF);
  float* sDelta = sLse + 128;
  uint64_t* bars = (uint64_t*)(smem + L::BAR_OFF);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2);

  const int p = blockIdx.x, h = blockIdx.y;
  const int s0 = cu[p], len = cu[p + 1] - s0;
  if (len <= 0) return;
  const int t = threadIdx.x, r = t & 127, half = t >> 7, warp = t >> 5;
  if (t == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);

  const bool live = r < len;
  const int pt = live ? gather_idx[s0 + r] : 0;
  const bool real = live && slot_of_point[pt] == s0 + r;   // alias slots are keys only: their dO (and so dQ) is zero
  const int nchunk = Dh >> 3;
  if (half == 0) {
    uint4 dov[DHP / 8];
    stage_row<DHP>((const uint4*)(q + (int64_t)pt * ldq + h * Dh), live, nchunk, r, sQ, sQt, nullptr);
    stage_row<DHP>((const uint4*)(dout + (int64_t)pt * ldo + h * Dh), real, nchunk, r, sdO, sdOt, dov);
    float dsum = 0.f;
    if (real) {
      const uint4* go = (const uint4*)(out + (int64_t)pt * ldo + h * Dh);
#pragma unroll
      for (int c = 0; c < DHP / 8; ++c) {
        if (c < nchunk) {
          uint4 ov = go[c];
          const __nv_bfloat162* a2 = (const __nv_bfloat162*)&dov[c];
          const __nv_bfloat162* b2 = (const __nv_bfloat162*)&ov;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 x = __bfloat1622float2(a2[e]), y = __bfloat1622float2(b2[e]);
            dsum = fmaf(x.x, y.x, fmaf(x.y, y.y, dsum));
          }
        }
      }
    }
    sDelta[r] = dsum;
    sLse[r] = live ? lse[(int64_t)(s0 + r) * H + h] : INFINITY;   // dead query columns: exp(-inf) = 0
  } else {
    stage_row<DHP>((const uint4*)(k + (int64_t)pt * ldk + h * Dh), live, nchunk, r, sK, sKt, nullptr);
    stage_row<DHP>((const uint4*)(v + (int64_t)pt * ldv + h * Dh), live, nchunk, r, sV, nullptr, nullptr);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- phase A: S^T = K Q^T (cols 0-127), dP^T = V dO^T (cols 128-255)
  if (t == 0) {
    const uint32_t idesc = make_idesc(128, 128, false, false);
    const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK), va = smem_u32(sV), da = smem_u32(sdO);
#pragma unroll
    for (int kk = 0; kk < DHP / 16; ++kk)
      umma_bf16(tmem_base, make_desc(ka + kk * 32, 16, 1024), make_desc(qa + kk * 32, 16, 1024), idesc, kk > 0 ? 1u : 0u);
#pragma unroll
    for (int kk = 0; kk < DHP / 16; ++kk)
      umma_bf16(tmem_base + 128, make_desc(va + kk * 32, 16, 1024), make_desc(da + kk * 32, 16, 1024), idesc, kk > 0 ? 1u : 0u);
    umma_commit(&bars[0]);
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after();

  // ---- thread (key r, q-half): P^T and dS^T for queries half*64 .. +63
  {
    float sv[64], dpv[64];
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + half * 64;
    tmem_ld_32x32b_x32(lane_addr, (uint32_t*)&sv[0]);
    tmem_ld_32x32b_x32(lane_addr + 32, (uint32_t*)&sv[32]);
    tmem_ld_32x32b_x32(lane_addr + 128, (uint32_t*)&dpv[0]);
    tmem_ld_32x32b_x32(lane_addr + 160, (uint32_t*)&dpv[32]);
    tmem_ld_wait();
    const float ks = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const uint32_t sid = layer_id * 0x9E37u + (uint32_t)(p * H + h);
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      const int qi = half * 64 + j;
      float pij = live ? __expf(sv[j] * scale - sLse[qi]) : 0.f;
      float pd = pij, dp = dpv[j];
      if (drop_p > 0.f) {
        bool keep = dropout_keep(seed, sid, (uint32_t)(qi * len + r), drop_p);
        pd = keep ? pij * ks : 0.f;
        dp = keep ? dp * ks : 0.f;
      }
      sv[j] = pd;
      dpv[j] = pij * (dp - sDelta[qi]) * scale;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint4 w, u;
      w.x = pack_bf16x2(sv[c * 8 + 0], sv[c * 8 + 1]);
      w.y = pack_bf16x2(sv[c * 8 + 2], sv[c * 8 + 3]);
      w.z = pack_bf16x2(sv[c * 8 + 4], sv[c * 8 + 5]);
      w.w = pack_bf16x2(sv[c * 8 + 6], sv[c * 8 + 7]);
      u.x = pack_bf16x2(dpv[c * 8 + 0], dpv[c * 8 + 1]);
      u.y = pack_bf16x2(dpv[c * 8 + 2], dpv[c * 8 + 3]);
      u.z = pack_bf16x2(dpv[c * 8 + 4], dpv[c * 8 + 5]);
      u.w = pack_bf16x2(dpv[c * 8 + 6], dpv[c * 8 + 7]);
      const int off = half * 16384 + r * 128 + ((c ^ (r & 7)) << 4);
      *(uint4*)(sPt + off) = w;
      *(uint4*)(sdSt + off) = u;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- phase B: dV = P^T dO (cols 0..), dK = dS^T Q (cols 64..), dQ = dS K (cols 128.., A = dS^T read MN-major)
  if (t == 0) {
    const uint32_t idesc = make_idesc(128, DHP, false, false);
    const uint32_t idesc_mn = make_idesc(128, DHP, true, false);
    const uint32_t pa = smem_u32(sPt), sa = smem_u32(sdSt), qt = smem_u32(sQt), kt = smem_u32(sKt), dt = smem_u32(sdOt);
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const uint32_t ao = (kk >> 2) * 16384 + (kk & 3) * 32, bo = (kk >> 2) * (DHP * 128) + (kk & 3) * 32;
      umma_bf16(tmem_base, make_desc(pa + ao, 16, 1024), make_desc(dt + bo, 16, 1024), idesc, kk > 0 ? 1u : 0u);
    }
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const uint32_t ao = (kk >> 2) * 16384 + (kk & 3) * 32, bo = (kk >> 2) * (DHP * 128) + (kk & 3) * 32;
      umma_bf16(tmem_base + 64, make_desc(sa + ao, 16, 1024), make_desc(qt + bo, 16, 1024), idesc, kk > 0 ? 1u : 0u);
    }
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const uint32_t bo = (kk >> 2) * (DHP * 128) + (kk & 3) * 32;
      umma_bf16(tmem_base + 128, make_desc(sa + kk * 2048, 16384, 1024), make_desc(kt + bo, 16, 1024), idesc_mn, kk > 0 ? 1u : 0u);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  {
    // warps 0-3: dV and dK of key r; warps 4-7: dQ of query r
    constexpr int NLD = (DHP + 31) / 32;
    float a[NLD * 32], b[NLD * 32];
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    if (half == 0) {
#pragma unroll
      for (int c = 0; c < NLD; ++c) {
        tmem_ld_32x32b_x32(lane_addr + c * 32, (uint32_t*)&a[c * 32]);
        tmem_ld_32x32b_x32(lane_addr + 64 + c * 32, (uint32_t*)&b[c * 32]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < NLD; ++c) tmem_ld_32x32b_x32(lane_addr + 128 + c * 32, (uint32_t*)&a[c * 32]);
    }
    tmem_ld_wait();
    if (half == 0 && live) {
      float* gv = dv + (int64_t)pt * lddv + h * Dh;
      float* gk = dk + (int64_t)pt * lddk + h * Dh;
#pragma unroll
      for (int c = 0; c < DHP / 4; ++c) {
        if (c * 4 < Dh) {
          red_add_v4(gv + c * 4, a[c * 4], a[c * 4 + 1], a[c * 4 + 2], a[c * 4 + 3]);
          red_add_v4(gk + c * 4, b[c * 4], b[c * 4 + 1], b[c * 4 + 2], b[c * 4 + 3]);
        }
      }
    } else if (half == 1 && real) {
      float* gq = dq + (int64_t)pt * lddq + h * Dh;
#pragma unroll
      for (int c = 0; c < DHP / 4; ++c)
        if (c * 4 < Dh) red_add_v4(gq + c * 4, a[c * 4], a[c * 4 + 1], a[c * 4 + 2], a[c * 4 + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

template <int DHP>
static int launch_bwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, const int32_t* gather_idx,
                      const int32_t* slot_of_point, const int32_t* cu, int n_patches, int H, int Dh, float scale, const void* out,
                      const void* dout, int64_t ldo, const float* lse, float* dq, float* dk, float* dv, int64_t lddq, int64_t lddk,
                      int64_t lddv, float drop_p, uint64_t seed, uint32_t layer_id, cudaStream_t st) {
  constexpr int SMEM = BwdSmem<DHP>::TOTAL;
  static bool attr_set = false;
  if (!attr_set) {
    LOTUS_CUDA_CALL(cudaFuncSetAttribute(k_patch_attn_bwd_tc<DHP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  k_patch_attn_bwd_tc<DHP><<<dim3(n_patches, H), 256, SMEM, st>>>(
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ldq, ldk, ldv, gather_idx, slot_of_point, cu, H, Dh,
      scale, (const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, ldo, lse, dq, dk, dv, lddq, lddk, lddv, drop_p, seed, layer_id);
  LOTUS_CUDA_OK();
  return 0;
}

}  // namespace tca

// Whether the tensor-core patch attention covers this problem (else the SIMT kernel of attention.cu runs):
// patches of at most 128 tokens, head dim a multiple of 8 up to 64, 16-byte aligned rows, no RPE.
bool lotus_patch_attn_tc_ok(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, const void* out,
                            int64_t ldo, int max_len, int Dh, const float* rpe_table) {
  auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  return rpe_table == nullptr && max_len <= 128 && Dh % 8 == 0 && Dh >= 8 && Dh <= 64 && al(q) && al(k) && al(v) && al(out) &&
         ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && getenv("LOTUS_ATTN_SIMT") == nullptr;
}

int lotus_patch_attn_fwd_tc(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                            const int32_t* gather_idx, const int32_t* slot_of_point, const int32_t* cu_seqlens, int n_patches, int H,
                            int Dh, float scale, void* out, int64_t ldo, float* lse, float drop_p, uint64_t seed, uint32_t layer_id,
                            void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_patches <= 0) return 0;
  if (Dh <= 16)
    return tca::launch_fwd<16>(q, k, v, ldq, ldk, ldv, gather_idx, slot_of_point, cu_seqlens, n_patches, H, Dh, scale, out, ldo, lse,
                               drop_p, seed, layer_id, st);
  if (Dh <= 32)
    return tca::launch_fwd<32>(q, k, v, ldq, ldk, ldv, gather_idx, slot_of_point, cu_seqlens, n_patches, H, Dh, scale, out, ldo, lse,
                               drop_p, seed, layer_id, st);
  return tca::launch_fwd<64>(q, k, v, ldq, ldk, ldv, gather_idx, slot_of_point, cu_seqlens, n_patches, H, Dh, scale, out, ldo, lse,
                             drop_p, seed, layer_id, st);
}

// dq / dk / dv rows must be 16-byte aligned for the vector atomics (red.global.add.v4.f32)
bool lotus_patch_attn_bwd_tc_ok(const float* dq, const float* dk, const float* dv, int64_t lddq, int64_t lddk, int64_t lddv, int Dh,
                                const float* drpe_table) {
  auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  return drpe_table == nullptr && Dh % 4 == 0 && al(dq) && al(dk) && al(dv) && lddq % 4 == 0 && lddk % 4 == 0 && lddv % 4 == 0;
}

int lotus_patch_attn_bwd_tc(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                            const int32_t* gather_idx, const int32_t* slot_of_point, const int32_t* cu_seqlens, int n_patches, int H,
                            int Dh, float scale, const void* out, const void* dout, int64_t ldo, const float* lse, float* dq, float* dk,
                            float* dv, int64_t lddq, int64_t lddk, int64_t lddv, float drop_p, uint64_t seed, uint32_t layer_id,
                            void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_patches <= 0) return 0;
  if (Dh <= 16)
    return tca::launch_bwd<16>(q, k, v, ldq, ldk, ldv, gather_idx, slot_of_point, cu_seqlens, n_patches, H, Dh, scale, out, dout, ldo,
                               lse, dq, dk, dv, lddq, lddk, lddv, drop_p, seed, layer_id, st);
  if (Dh <= 32)
    return tca::launch_bwd<32>(q, k, v, ldq, ldk, ldv, gather_idx, slot_of_point, cu_seqlens, n_patches, H, Dh, scale, out, dout, ldo,
                               lse, dq, dk, dv, lddq, lddk, lddv, drop_p, seed, layer_id, st);
  return tca::launch_bwd<64>(q, k, v, ldq, ldk, ldv, gather_idx, slot_of_point, cu_seqlens, n_patches, H, Dh, scale, out, dout, ldo,
                             lse, dq, dk, dv, lddq, lddk, lddv, drop_p, seed, layer_id, st);
}
