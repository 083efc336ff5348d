// Cross attention of packed points to their cloud's CLIP tokens, bf16 tensor-core mode
// (reference: models/PointTransformerV3/model_ca.py:46-101; the fp32 validation-mode kernels are in attention.cu).
//
// The op is HBM-bound by a wide margin (L <= 32 tokens: ~20 FLOP per byte of q/out traffic), so the contractions use
// warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate) straight from registers: a tcgen05 tile (128 x N x 16, TMEM
// round trip) has nothing to amortise at N = L.  One warp owns 16 points x one head; K / V of the (cloud, head) sit in
// shared memory as bf16; tokens are walked in chunks of 32, so any L <= 128 works.
//   forward : S = q K^T * scale -> row max (pass 1) -> P = exp(S - m), dropout, O = P V (pass 2) -> out, lse
//   backward: P = exp(S - lse); dP = dO V^T; dS = P * (mask * dP / keep - D); dQ = dS K * scale;
//             dK = scale * dS^T q, dV = Pd^T dO via register transposes (movmatrix) of the accumulator fragments,
//             reduced over the CTA in shared memory, then ONE global atomicAdd per (token, channel) per CTA.
// The dropout mask is the same counter hash as the SIMT kernels (element (point * 128 + token), stream = layer, head).
#include "common.cuh"

namespace xattn {

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// transpose an 8x8 b16 tile held one 32-bit register per lane (row = lane / 4, columns 2 * (lane % 4) + {0, 1})
__device__ __forceinline__ uint32_t movm_trans(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;\n" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// K and V of one (cloud, head) -> shared memory, bf16, rows padded to a multiple of 32 tokens and DHP (+8) columns
template <int DHP>
__device__ __forceinline__ void load_kv(const bf16_t* __restrict__ k, const bf16_t* __restrict__ v, int64_t ldk, int64_t ldv, int ts, int L,
                                        int LPAD, int h, int DH, uint32_t* Ks, uint32_t* Vs) {
  constexpr int W = (DHP + 8) / 2;          // row stride in 32-bit words
  for (int t = threadIdx.x; t < LPAD * (DHP / 2); t += blockDim.x) {
    int row = t / (DHP / 2), cp = t - row * (DHP / 2), col = 2 * cp;
    uint32_t kk = 0, vv = 0;
    if (row < L && col < DH) {
      kk = *reinterpret_cast<const uint32_t*>(k + (int64_t)(ts + row) * ldk + h * DH + col);
      vv = *reinterpret_cast<const uint32_t*>(v + (int64_t)(ts + row) * ldv + h * DH + col);
    }
    Ks[row * W + cp] = kk;
    Vs[row * W + cp] = vv;
  }
}
// A fragments (16 rows x DHP) of a row-major bf16 matrix straight from global memory
template <int DHP>
__device__ __forceinline__ void load_a(const bf16_t* __restrict__ x, int64_t ld, int64_t r0, int64_t r1, bool v0, bool v1, int h, int DH, int t,
                                       uint32_t (&a)[DHP / 16][4]) {
#pragma unroll
  for (int ks = 0; ks < DHP / 16; ++ks) {
    int c0 = 16 * ks + 2 * t, c1 = c0 + 8;
    a[ks][0] = (v0 && c0 < DH) ? *reinterpret_cast<const uint32_t*>(x + r0 * ld + h * DH + c0) : 0u;
    a[ks][1] = (v1 && c0 < DH) ? *reinterpret_cast<const uint32_t*>(x + r1 * ld + h * DH + c0) : 0u;
    a[ks][2] = (v0 && c1 < DH) ? *reinterpret_cast<const uint32_t*>(x + r0 * ld + h * DH + c1) : 0u;
    a[ks][3] = (v1 && c1 < DH) ? *reinterpret_cast<const uint32_t*>(x + r1 * ld + h * DH + c1) : 0u;
  }
}
// C[16 x 32 tokens] = A[16 x DHP] * M[tokens c0.. x DHP]^T  (M = K for the scores, V for dP)
template <int DHP>
__device__ __forceinline__ void mm_rows_by_tokens(const uint32_t (&a)[DHP / 16][4], const uint32_t* Ms, int tok0, int g, int t, float (&c)[4][4]) {
  constexpr int W = (DHP + 8) / 2;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
    const uint32_t* row = Ms + (tok0 + 8 * nt + g) * W;
#pragma unroll
    for (int ks = 0; ks < DHP / 16; ++ks) mma16816(c[nt], a[ks], row[8 * ks + t], row[8 * ks + 4 + t]);
  }
}
// acc[16 x DHP] += P[16 x 32 tokens] * M[tokens tok0.. x DHP]   (M = V for the output, K for dQ); p: packed bf16 A fragments
template <int DHP>
__device__ __forceinline__ void mm_tokens_by_cols(const uint32_t (&p)[2][4], const uint32_t* Ms, int tok0, int g, int t, float (&acc)[DHP / 8][4]) {
  constexpr int LDB = DHP + 8;              // row stride in bf16 elements
  const uint16_t* M16 = reinterpret_cast<const uint16_t*>(Ms);
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {
    int r = tok0 + 16 * kb + 2 * t;
#pragma unroll
    for (int dn = 0; dn < DHP / 8; ++dn) {
      int col = 8 * dn + g;
      uint32_t b0 = (uint32_t)M16[r * LDB + col] | ((uint32_t)M16[(r + 1) * LDB + col] << 16);
      uint32_t b1 = (uint32_t)M16[(r + 8) * LDB + col] | ((uint32_t)M16[(r + 9) * LDB + col] << 16);
      mma16816(acc[dn], p[kb], b0, b1);
    }
  }
}

template <int DHP>
__global__ void __launch_bounds__(128) k_cross_fwd_mma(
    const bf16_t* __restrict__ q, int64_t ldq, const bf16_t* __restrict__ k, const bf16_t* __restrict__ v, int64_t ldk, int64_t ldv,
    const int32_t* __restrict__ pt_off, const int32_t* __restrict__ tok_off, int H, int DH, float scale, bf16_t* __restrict__ out,
    int64_t ldo, float* __restrict__ lse, float drop_p, uint64_t seed, uint32_t layer_id) {
  extern __shared__ __align__(16) uint32_t smw[];
  constexpr int W = (DHP + 8) / 2;
  const int b = blockIdx.y, h = blockIdx.z;
  const int ps = pt_off[b], pn = pt_off[b + 1] - ps;
  const int i0 = blockIdx.x * 128;
  if (i0 >= pn) return;
  const int ts = tok_off[b], L = tok_off[b + 1] - ts;
  const int LPAD = (L + 31) & ~31;
  uint32_t* Ks = smw;
  uint32_t* Vs = Ks + LPAD * W;
  load_kv<DHP>(k, v, ldk, ldv, ts, L, LPAD, h, DH, Ks, Vs);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int cn = min(128, pn - i0);
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t sid = layer_id * 0x9E37u + (uint32_t)h;
  for (int mt = warp; mt * 16 < cn; mt += 4) {
    const int r0 = i0 + mt * 16 + g, r1 = r0 + 8;
    const bool v0 = r0 < pn, v1 = r1 < pn;
    const int64_t p0 = ps + r0, p1 = ps + r1;
    uint32_t aq[DHP / 16][4];
    load_a<DHP>(q, ldq, p0, p1, v0, v1, h, DH, t, aq);
    float s[4][4];
    // pass 1: row maxima over all tokens
    float m0 = -INFINITY, m1 = -INFINITY;
    for (int c0 = 0; c0 < LPAD; c0 += 32) {
      mm_rows_by_tokens<DHP>(aq, Ks, c0, g, t, s);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        int j = c0 + 8 * nt + 2 * t;
        if (j < L) { m0 = fmaxf(m0, s[nt][0] * scale); m1 = fmaxf(m1, s[nt][2] * scale); }
        if (j + 1 < L) { m0 = fmaxf(m0, s[nt][1] * scale); m1 = fmaxf(m1, s[nt][3] * scale); }
      }
    }
    m0 = quad_max(m0); m1 = quad_max(m1);
    // pass 2: probabilities, dropout, O = P V
    float o[DHP / 8][4];
#pragma unroll
    for (int dn = 0; dn < DHP / 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
    float l0 = 0.f, l1 = 0.f;
    for (int c0 = 0; c0 < LPAD; c0 += 32) {
      mm_rows_by_tokens<DHP>(aq, Ks, c0, g, t, s);
      uint32_t pa[2][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int j = c0 + 8 * nt + 2 * t + (i & 1);
          float mm = (i < 2) ? m0 : m1;
          float pj = (j < L) ? __expf(s[nt][i] * scale - mm) : 0.f;
          if (i < 2) l0 += pj; else l1 += pj;
          if (drop_p > 0.f && j < L) {
            uint32_t pt = (uint32_t)((i < 2) ? p0 : p1);
            pj = dropout_keep(seed, sid, pt * 128u + (uint32_t)j, drop_p) ? pj * keep_scale : 0.f;
          }
          e[i] = pj;
        }
        pa[nt >> 1][(nt & 1) * 2 + 0] = pack2(e[0], e[1]);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack2(e[2], e[3]);
      }
      mm_tokens_by_cols<DHP>(pa, Vs, c0, g, t, o);
    }
    l0 = quad_sum(l0); l1 = quad_sum(l1);
    const float inv0 = L > 0 ? 1.f / l0 : 0.f, inv1 = L > 0 ? 1.f / l1 : 0.f;
    if (t == 0) {
      if (v0) lse[p0 * H + h] = m0 + __logf(l0);
      if (v1) lse[p1 * H + h] = m1 + __logf(l1);
    }
#pragma unroll
    for (int dn = 0; dn < DHP / 8; ++dn) {
      int col = 8 * dn + 2 * t;
      if (col < DH) {
        if (v0) *reinterpret_cast<uint32_t*>(out + p0 * ldo + h * DH + col) = pack2(o[dn][0] * inv0, o[dn][1] * inv0);
        if (v1) *reinterpret_cast<uint32_t*>(out + p1 * ldo + h * DH + col) = pack2(o[dn][2] * inv1, o[dn][3] * inv1);
      }
    }
  }
}

template <int DHP>
__global__ void __launch_bounds__(128) k_cross_bwd_mma(
    const bf16_t* __restrict__ q, int64_t ldq, const bf16_t* __restrict__ k, const bf16_t* __restrict__ v, int64_t ldk, int64_t ldv,
    const int32_t* __restrict__ pt_off, const int32_t* __restrict__ tok_off, int H, int DH, float scale, const bf16_t* __restrict__ out,
    const bf16_t* __restrict__ dout, int64_t ldo, const float* __restrict__ lse, float* __restrict__ dq, int64_t lddq,
    float* __restrict__ dk, float* __restrict__ dv, int64_t lddk, int64_t lddv, float drop_p, uint64_t seed, uint32_t layer_id) {
  extern __shared__ __align__(16) uint32_t smw[];
  constexpr int W = (DHP + 8) / 2;
  const int b = blockIdx.y, h = blockIdx.z;
  const int ps = pt_off[b], pn = pt_off[b + 1] - ps;
  const int i0 = blockIdx.x * 128;
  if (i0 >= pn) return;
  const int ts = tok_off[b], L = tok_off[b + 1] - ts;
  const int LPAD = (L + 31) & ~31;
  uint32_t* Ks = smw;
  uint32_t* Vs = Ks + LPAD * W;
  float* dKs = reinterpret_cast<float*>(Vs + LPAD * W);      // [LPAD][DHP] fp32, CTA-wide accumulators
  float* dVs = dKs + LPAD * DHP;
  load_kv<DHP>(k, v, ldk, ldv, ts, L, LPAD, h, DH, Ks, Vs);
  for (int i = threadIdx.x; i < 2 * LPAD * DHP; i += blockDim.x) dKs[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int cn = min(128, pn - i0);
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t sid = layer_id * 0x9E37u + (uint32_t)h;
  const bool one_chunk = LPAD == 32;
  float ak[2][DHP / 8][4], av[2][DHP / 8][4];                 // dK / dV of the current 32-token chunk (tokens x channels)
  auto clear_kv = [&]() {
#pragma unroll
    for (int m2 = 0; m2 < 2; ++m2)
#pragma unroll
      for (int dn = 0; dn < DHP / 8; ++dn)
#pragma unroll
        for (int i = 0; i < 4; ++i) { ak[m2][dn][i] = 0.f; av[m2][dn][i] = 0.f; }
  };
  auto flush_kv = [&](int c0) {
#pragma unroll
    for (int m2 = 0; m2 < 2; ++m2)
#pragma unroll
      for (int dn = 0; dn < DHP / 8; ++dn)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int tok = c0 + 16 * m2 + g + ((i >> 1) ? 8 : 0), col = 8 * dn + 2 * t + (i & 1);
          atomicAdd(&dKs[tok * DHP + col], ak[m2][dn][i]);
          atomicAdd(&dVs[tok * DHP + col], av[m2][dn][i]);
        }
  };
  clear_kv();
  for (int mt = warp; mt * 16 < cn; mt += 4) {
    const int r0 = i0 + mt * 16 + g, r1 = r0 + 8;
    const bool v0 = r0 < pn, v1 = r1 < pn;
    const int64_t p0 = ps + r0, p1 = ps + r1;
    uint32_t aq[DHP / 16][4], ad[DHP / 16][4], tq[DHP / 16][4], td[DHP / 16][4];
    load_a<DHP>(q, ldq, p0, p1, v0, v1, h, DH, t, aq);
    load_a<DHP>(dout, ldo, p0, p1, v0, v1, h, DH, t, ad);
    // D = rowsum(dO * O)
    float d0 = 0.f, d1 = 0.f;
    {
      uint32_t ao[DHP / 16][4];
      load_a<DHP>(out, ldo, p0, p1, v0, v1, h, DH, t, ao);
#pragma unroll
      for (int ks = 0; ks < DHP / 16; ++ks)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 x = unpack2(ad[ks][i]), y = unpack2(ao[ks][i]);
          float pr = x.x * y.x + x.y * y.y;
          if (i & 1) d1 += pr; else d0 += pr;
        }
    }
    d0 = quad_sum(d0); d1 = quad_sum(d1);
#pragma unroll
    for (int ks = 0; ks < DHP / 16; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i) { tq[ks][i] = movm_trans(aq[ks][i]); td[ks][i] = movm_trans(ad[ks][i]); }
    const float ls0 = v0 ? lse[p0 * H + h] : 0.f, ls1 = v1 ? lse[p1 * H + h] : 0.f;
    float dqa[DHP / 8][4];
#pragma unroll
    for (int dn = 0; dn < DHP / 8; ++dn) dqa[dn][0] = dqa[dn][1] = dqa[dn][2] = dqa[dn][3] = 0.f;
    for (int c0 = 0; c0 < LPAD; c0 += 32) {
      float s[4][4], dp[4][4];
      mm_rows_by_tokens<DHP>(aq, Ks, c0, g, t, s);
      mm_rows_by_tokens<DHP>(ad, Vs, c0, g, t, dp);
      uint32_t dsa[2][4];        // dS as A fragments (points x tokens) for dQ
      uint32_t dst[2][4];        // dS^T as A fragments (tokens x points) for dK
      uint32_t pdt[2][4];        // Pd^T as A fragments for dV
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float ds[4], pd[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int j = c0 + 8 * nt + 2 * t + (i & 1);
          bool rv = (i < 2) ? v0 : v1;
          float pij = (j < L && rv) ? __expf(s[nt][i] * scale - ((i < 2) ? ls0 : ls1)) : 0.f;
          float dpi = dp[nt][i], pdi = pij;
          if (drop_p > 0.f && j < L) {
            uint32_t pt = (uint32_t)((i < 2) ? p0 : p1);
            bool keep = dropout_keep(seed, sid, pt * 128u + (uint32_t)j, drop_p);
            pdi = keep ? pij * keep_scale : 0.f;
            dpi = keep ? dpi * keep_scale : 0.f;
          }
          ds[i] = pij * (dpi - ((i < 2) ? d0 : d1));
          pd[i] = pdi;
        }
        uint32_t ds01 = pack2(ds[0], ds[1]), ds23 = pack2(ds[2], ds[3]);
        uint32_t pd01 = pack2(pd[0], pd[1]), pd23 = pack2(pd[2], pd[3]);
        dsa[nt >> 1][(nt & 1) * 2 + 0] = ds01;
        dsa[nt >> 1][(nt & 1) * 2 + 1] = ds23;
        // token tile m2 = nt / 2: a0 = T(rows 0-7, tokens 16 m2 ..+7), a1 = T(rows 0-7, tokens +8..), a2 / a3 = rows 8-15
        dst[nt >> 1][(nt & 1) + 0] = movm_trans(ds01);
        dst[nt >> 1][(nt & 1) + 2] = movm_trans(ds23);
        pdt[nt >> 1][(nt & 1) + 0] = movm_trans(pd01);
        pdt[nt >> 1][(nt & 1) + 2] = movm_trans(pd23);
      }
      mm_tokens_by_cols<DHP>(dsa, Ks, c0, g, t, dqa);
      // dK[tokens x channels] += dS^T q ; dV += Pd^T dO   (B fragments = register transposes of the q / dO A fragments)
#pragma unroll
      for (int m2 = 0; m2 < 2; ++m2)
#pragma unroll
        for (int dn = 0; dn < DHP / 8; ++dn) {
          int ks = dn >> 1, hf = (dn & 1) * 2;
          mma16816(ak[m2][dn], dst[m2], tq[ks][hf], tq[ks][hf + 1]);
          mma16816(av[m2][dn], pdt[m2], td[ks][hf], td[ks][hf + 1]);
        }
      if (!one_chunk) { flush_kv(c0); clear_kv(); }
    }
#pragma unroll
    for (int dn = 0; dn < DHP / 8; ++dn) {
      int col = 8 * dn + 2 * t;
      if (col < DH) {
        if (v0) *reinterpret_cast<float2*>(dq + p0 * lddq + h * DH + col) = make_float2(dqa[dn][0] * scale, dqa[dn][1] * scale);
        if (v1) *reinterpret_cast<float2*>(dq + p1 * lddq + h * DH + col) = make_float2(dqa[dn][2] * scale, dqa[dn][3] * scale);
      }
    }
  }
  if (one_chunk) flush_kv(0);
  __syncthreads();
  for (int o = threadIdx.x; o < L * DH; o += blockDim.x) {
    int j = o / DH, d = o - j * DH;
    atomicAdd(&dk[(int64_t)(ts + j) * lddk + h * DH + d], dKs[j * DHP + d] * scale);
    atomicAdd(&dv[(int64_t)(ts + j) * lddv + h * DH + d], dVs[j * DHP + d]);
  }
}

}  // namespace xattn

#define XATTN_DISPATCH(DHV, ...)                                    \
  if ((DHV) <= 16) { constexpr int DHP = 16; __VA_ARGS__; }          \
  else if ((DHV) <= 32) { constexpr int DHP = 32; __VA_ARGS__; }     \
  else { constexpr int DHP = 64; __VA_ARGS__; }

// the mma path needs 4-byte aligned bf16 pairs / 8-byte aligned fp32 pairs and an even head width
bool lotus_cross_attn_mma_ok(const void* q, const void* k, const void* v, const void* out, int64_t ldq, int64_t ldk, int64_t ldv,
                             int64_t ldo, int Dh, int max_tokens) {
  auto al4 = [](const void* p) { return ((uintptr_t)p & 3) == 0; };
  return Dh % 2 == 0 && Dh >= 8 && Dh <= 64 && max_tokens <= 128 && al4(q) && al4(k) && al4(v) && al4(out) && ldq % 2 == 0 &&
         ldk % 2 == 0 && ldv % 2 == 0 && ldo % 2 == 0 && getenv("LOTUS_CROSS_SIMT") == nullptr;
}

int lotus_cross_attn_fwd_mma(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldk, int64_t ldv, const int32_t* pt_off,
                             const int32_t* tok_off, int B, int max_points, int max_tokens, int H, int Dh, float scale, void* out,
                             int64_t ldo, float* lse, float drop_p, uint64_t seed, uint32_t layer_id, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (B <= 0 || max_points <= 0) return 0;
  int LPAD = (max_tokens + 31) & ~31;
  dim3 grid(ceil_div(max_points, 128), B, H);
  XATTN_DISPATCH(Dh, {
    size_t smem = (size_t)2 * LPAD * (DHP + 8) * 2;
    xattn::k_cross_fwd_mma<DHP><<<grid, 128, smem, st>>>((const bf16_t*)q, ldq, (const bf16_t*)k, (const bf16_t*)v, ldk, ldv, pt_off, tok_off,
                                                         H, Dh, scale, (bf16_t*)out, ldo, lse, drop_p, seed, layer_id);
  });
  LOTUS_CUDA_OK();
  return 0;
}

// dq (fp32) is written; dk / dv (fp32) are ACCUMULATED with atomics: zero them first.
int lotus_cross_attn_bwd_mma(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldk, int64_t ldv, const int32_t* pt_off,
                             const int32_t* tok_off, int B, int max_points, int max_tokens, int H, int Dh, float scale, const void* out,
                             const void* dout, int64_t ldo, const float* lse, float* dq, int64_t lddq, float* dk, float* dv,
                             int64_t lddk, int64_t lddv, float drop_p, uint64_t seed, uint32_t layer_id, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (B <= 0 || max_points <= 0) return 0;
  int LPAD = (max_tokens + 31) & ~31;
  dim3 grid(ceil_div(max_points, 128), B, H);
  XATTN_DISPATCH(Dh, {
    size_t smem = (size_t)2 * LPAD * (DHP + 8) * 2 + (size_t)2 * LPAD * DHP * 4;
    LOTUS_CUDA_CALL(cudaFuncSetAttribute(xattn::k_cross_bwd_mma<DHP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    xattn::k_cross_bwd_mma<DHP><<<grid, 128, smem, st>>>((const bf16_t*)q, ldq, (const bf16_t*)k, (const bf16_t*)v, ldk, ldv, pt_off, tok_off,
                                                         H, Dh, scale, (const bf16_t*)out, (const bf16_t*)dout, ldo, lse, dq, lddq, dk, dv,
                                                         lddk, lddv, drop_p, seed, layer_id);
  });
  LOTUS_CUDA_OK();
  return 0;
}
