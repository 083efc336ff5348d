// Fused patch self-attention for sm_100a (bf16 activations): SerializedAttention.forward of the reference
// (models/PointTransformerV3/model.py:468-557, flash branch; RPE of the non-flash branch :307-326, 400-408, 520-521) in ONE kernel
// per direction:
//
//   gather q/k/v rows by order[pad]  ->  per-head LayerNorm of q and k (model.py:393-394, 532-535; eps 1e-6) in registers
//   ->  S = Q K^T on tcgen05 (TMEM accumulator)  ->  scale, + relative-position bias, softmax, dropout in the epilogue of that MMA
//   ->  O = P V on tcgen05  ->  normalise, scatter by unpad[inverse] (alias tail slots are keys only).
//
// What changed against attention_tc.cu (round 1; one CTA per (patch, head), q/k normalised by separate launches):
//   * one CTA per (patch, PAIR of heads): a gathered row is read as one full 128-byte line per tensor and one smem row serves both heads
//     (head j lives at byte offset j * 2 * DHP of the row, which is a legal K-offset inside the SWIZZLE_128B atom);
//   * the per-head q / k LayerNorm runs on the gathered row in registers: 36 LayerNorm launches and the qh / kh tensors (one write and
//     one read of q and k per block) are gone;
//   * V is consumed MN-major straight from its natural [key][d] rows -- no transposed copy (32 scalar shared stores per thread before);
//   * softmax makes two passes over TMEM (max; exp / sum / pack) instead of holding 128 logits per thread: ~60 registers, two CTAs/SM;
//   * the RPE bias (table lookup on relative grid coordinates) is added to the logits in the same epilogue (north star: "per-point
//     position bias and softmax fused in the epilogue" of the MMA kernel).
//
// The backward (k_pattn_bwd) mirrors it: q / k re-normalised at load, S^T and dP^T on tcgen05, P^T / dS^T as K-major A operands,
// dV / dK / dQ against the natural row tiles as MN-major B operands (no transposed copies), per-head LayerNorm BACKWARD in the epilogue,
// and bf16 gradients written straight into the packed dqkv tensor; the second contribution of an alias tail slot goes to a side buffer
// that a small pass folds in, so the fp32 accumulators, their memsets, the vector atomics and the convert / LayerNorm-backward launches
// of round 1 disappear.
#include "tc_common.cuh"
#include <stdlib.h>

namespace paf {
using namespace tc;

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *(uint32_t*)&v;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int MAX_TAB = 288;     // RPE table rows held in shared memory (3 * (2 * pos_bnd + 1) <= 288: pos_bnd <= 47; K = 128 gives 15)

struct Args {
  const __nv_bfloat16 *q, *k, *v;           // raw projections (rows = points); ld in elements
  int64_t ldq, ldk, ldv;
  const float *qg, *qb, *kg, *kb;            // per-head LayerNorm affine [Dh] (null: no q/k normalisation)
  const int32_t *gather_idx, *slot_of_point, *cu;
  int H, Dh;
  float scale;
  const float* rpe_table;                    // [3 * (2 * pos_bnd + 1), H] or null
  const int32_t* grid;                       // [n, 3] grid coordinates (RPE)
  int pos_bnd;
  __nv_bfloat16* out;                        // forward: attention output; backward: the forward's output (read)
  int64_t ldo;
  float* lse;                                // [n_pad, H]
  float drop_p;
  uint64_t seed;
  uint32_t layer_id;
  // backward only
  const __nv_bfloat16* dout;
  __nv_bfloat16 *dq, *dk, *dv;               // gradients w.r.t. the RAW q / k / v rows (bf16), ld in elements
  int64_t lddq, lddk, lddv;
  __nv_bfloat16* alias_dkv;                  // [n_pad, 2, H * Dh] side buffer: dk / dv contributions of alias slots
  float *dqg, *dqb, *dkg, *dkb;              // LayerNorm affine gradients (accumulated)
  float* drpe;                               // RPE table gradient (accumulated) or null
  long long* dbg;                            // dev: per-CTA phase timestamps (8 x clock64 per CTA) or null
};
#define PTS(slot) do { if (a.dbg != nullptr && threadIdx.x == 0) a.dbg[(blockIdx.x + gridDim.x * blockIdx.y) * 8 + (slot)] = clock64(); } while (0)

// ------------------------------------------------------------------------------------------------ row staging
// One gathered row (HG heads x Dh values) -> optional per-head LayerNorm (fp32) -> bf16 -> K-major SWIZZLE_128B tile row r, head j at
// 16-byte chunks [j * DHP / 8, +DHP / 8) (zero padded).  `xh_out` (optional): the normalised values BEFORE the affine map, `rstd_out`
// the per-head 1/std (the backward's LayerNorm gradient needs both).
// the HG x DHP/8 16-byte chunks of one gathered row -> registers (zeros for dead rows, absent heads and the padding beyond Dh);
// issued BEFORE the barrier that publishes the LayerNorm parameters, so the row loads overlap it
template <int DHP, int HG>
__device__ __forceinline__ void load_row(const __nv_bfloat16* __restrict__ src, bool live, int nh, int Dh, uint4* regs) {
#pragma unroll
  for (int j = 0; j < HG; ++j)
#pragma unroll
    for (int c = 0; c < DHP / 8; ++c) {
      uint4 a = make_uint4(0, 0, 0, 0);
      if (live && j < nh && c * 8 < Dh) a = *(const uint4*)(src + j * Dh + c * 8);
      regs[j * (DHP / 8) + c] = a;
    }
}
// registers of one row -> optional per-head LayerNorm (fp32) -> bf16 -> K-major SWIZZLE_128B tile row r, head j at 16-byte chunks
// [j * DHP / 8, +DHP / 8) (zero padded)
template <int DHP, int HG>
__device__ __forceinline__ void ln_store_row(const uint4* regs, bool live, int nh, int Dh, const float* __restrict__ sg,
                                             const float* __restrict__ sb, int r, uint8_t* tile) {
#pragma unroll
  for (int j = 0; j < HG; ++j) {
    const bool on = live && j < nh;
    if (sg == nullptr || !on) {                                  // plain copy (no q/k normalisation, or an all-zero row)
#pragma unroll
      for (int c = 0; c < DHP / 8; ++c) *(uint4*)(tile + r * 128 + (((j * (DHP / 8) + c) ^ (r & 7)) << 4)) = regs[j * (DHP / 8) + c];
      continue;
    }
    float x[DHP];
#pragma unroll
    for (int c = 0; c < DHP / 8; ++c) {
      const __nv_bfloat162* h2 = (const __nv_bfloat162*)&regs[j * (DHP / 8) + c];
#pragma unroll
      for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(h2[e]); x[c * 8 + 2 * e] = f.x; x[c * 8 + 2 * e + 1] = f.y; }
    }
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < DHP; ++d) s += (d < Dh) ? x[d] : 0.f;
    const float mu = s / (float)Dh;
    float qv = 0.f;
#pragma unroll
    for (int d = 0; d < DHP; ++d) { float t = (d < Dh) ? x[d] - mu : 0.f; qv += t * t; }
    const float rs = rsqrtf(qv / (float)Dh + 1e-6f);
#pragma unroll
    for (int d = 0; d < DHP; ++d) x[d] = (d < Dh) ? (x[d] - mu) * rs * sg[d] + sb[d] : 0.f;
#pragma unroll
    for (int c = 0; c < DHP / 8; ++c) {
      uint4 w;
      w.x = pack2(x[c * 8 + 0], x[c * 8 + 1]); w.y = pack2(x[c * 8 + 2], x[c * 8 + 3]);
      w.z = pack2(x[c * 8 + 4], x[c * 8 + 5]); w.w = pack2(x[c * 8 + 6], x[c * 8 + 7]);
      *(uint4*)(tile + r * 128 + (((j * (DHP / 8) + c) ^ (r & 7)) << 4)) = w;
    }
  }
}
// plain copy of `nchunks` 16-byte chunks of a gathered row into chunks [c0, c0 + nchunks) of tile row r (zero when !on or beyond Dh)
template <int DHP>
__device__ __forceinline__ void stage_row_copy(const __nv_bfloat16* __restrict__ src, bool on, int Dh, int r, uint8_t* tile, int c0) {
#pragma unroll
  for (int c = 0; c < DHP / 8; ++c) {
    uint4 a = make_uint4(0, 0, 0, 0);
    if (on && c * 8 < Dh) a = *(const uint4*)(src + c * 8);
    *(uint4*)(tile + r * 128 + (((c0 + c) ^ (r & 7)) << 4)) = a;
  }
}


// ------------------------------------------------------------------------------------------------ coalesced gather / scatter
// The gathered tiles are 128 rows x 128 bytes.  With one row per lane every warp-level load touches 32 different lines = 32 L1TEX
// wavefronts, and the phase stamps (lotus_attn_debug_buffer) showed the kernels bound by exactly that.  Here LPR = HG*DHP/8 lanes
// cover ONE row (lane = 16-byte chunk), so a warp instruction touches 32/LPR lines; the per-head LayerNorm statistics are formed by
// shuffles inside the DHP/8 lanes that hold a head.
template <int DHP, int HG, int NT>
struct Coal {
  static constexpr int LPR = HG * DHP / 8;        // lanes per row
  static constexpr int RPP = NT / LPR;            // rows per pass of the CTA
  static constexpr int RPT = 128 / RPP;           // passes = rows per thread
  static constexpr int GL = DHP / 8;              // lanes per head
};
// source rows (points) of this thread's RPT tile rows; -1 for rows beyond the patch
template <int DHP, int HG, int NT>
__device__ __forceinline__ void coal_rows(const int32_t* __restrict__ gidx, int len, int t, int* pts) {
  using C = Coal<DHP, HG, NT>;
#pragma unroll
  for (int i = 0; i < C::RPT; ++i) { const int row = t / C::LPR + i * C::RPP; pts[i] = row < len ? gidx[row] : -1; }
}
// issue the loads of this thread's chunks of a gathered tensor (rows `pts`, columns col0 + head * Dh + chunk * 8) into registers
template <int DHP, int HG, int NT>
__device__ __forceinline__ void coal_load(const __nv_bfloat16* __restrict__ base, int64_t ld, int col0, const int* pts, int nh, int Dh, int t, uint4* regs) {
  using C = Coal<DHP, HG, NT>;
  const int ch = t % C::LPR, j = ch / C::GL, c = ch % C::GL;
#pragma unroll
  for (int i = 0; i < C::RPT; ++i) {
    regs[i] = make_uint4(0, 0, 0, 0);
    if (pts[i] >= 0 && j < nh && c * 8 < Dh) regs[i] = *(const uint4*)(base + (int64_t)pts[i] * ld + col0 + j * Dh + c * 8);
  }
}
// registers -> (optional per-head LayerNorm) -> K-major SWIZZLE_128B tile
template <int DHP, int HG, int NT>
__device__ __forceinline__ void coal_store(const uint4* regs, const int* pts, int nh, int Dh, const float* __restrict__ sg, const float* __restrict__ sb, int t,
                                           uint8_t* tile) {
  using C = Coal<DHP, HG, NT>;
  const int ch = t % C::LPR, j = ch / C::GL, c = ch % C::GL;
  const bool col_ok = j < nh && c * 8 < Dh;
#pragma unroll
  for (int i = 0; i < C::RPT; ++i) {
    const int row = t / C::LPR + i * C::RPP;
    uint4 w = regs[i];
    if (sg != nullptr) {
      const bool on = pts[i] >= 0 && col_ok;
      float x[8];
      const __nv_bfloat162* h2 = (const __nv_bfloat162*)&regs[i];
#pragma unroll
      for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(h2[e]); x[2 * e] = f.x; x[2 * e + 1] = f.y; }
      float sum = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += x[e];
#pragma unroll
      for (int o = C::GL >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mu = sum / (float)Dh;
      float q = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = on ? x[e] - mu : 0.f; q += d * d; }
#pragma unroll
      for (int o = C::GL >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rs = rsqrtf(q / (float)Dh + 1e-6f);
      float y[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = on ? (x[e] - mu) * rs * sg[c * 8 + e] + sb[c * 8 + e] : 0.f;
      w.x = pack2(y[0], y[1]); w.y = pack2(y[2], y[3]); w.z = pack2(y[4], y[5]); w.w = pack2(y[6], y[7]);
    }
    *(uint4*)(tile + row * 128 + ((ch ^ (row & 7)) << 4)) = w;
  }
}

// ------------------------------------------------------------------------------------------------ forward softmax passes
struct RowCtx {
  uint32_t lane_addr; int len, r, j; float sl2;
  const int* sGrid; const float* sTab; int bnd, num, g0, g1, g2;   // RPE: grid coordinates of this row, table slice [rows][HG]
  float ks; uint32_t hq, thr24;                                    // dropout: keep-scale, hash of this query row, 24-bit threshold
};
// logit of (this row, column jk) in the log2 domain; RPE: + table[clamp(g_row - g_col)] per axis (sign: row minus column)
template <int HG, bool RPE>
__device__ __forceinline__ float logit2(const RowCtx& c, uint32_t raw, int jk) {
  float x = __uint_as_float(raw) * c.sl2;
  if (RPE) {
    const int d0 = max(-c.bnd, min(c.bnd, c.g0 - c.sGrid[jk * 3])), d1 = max(-c.bnd, min(c.bnd, c.g1 - c.sGrid[jk * 3 + 1])),
              d2 = max(-c.bnd, min(c.bnd, c.g2 - c.sGrid[jk * 3 + 2]));
    x += (c.sTab[(d0 + c.bnd) * HG + c.j] + c.sTab[(d1 + c.bnd + c.num) * HG + c.j] + c.sTab[(d2 + c.bnd + 2 * c.num) * HG + c.j]) * 1.4426950408889634f;
  }
  return x;
}
template <int HG, bool RPE>
__device__ __forceinline__ float softmax_max(const RowCtx& c) {
  float m = -INFINITY;
#pragma unroll 1
  for (int cc = 0; cc < 8; ++cc) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(c.lane_addr + cc * 16, v);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int jk = cc * 16 + e;
      // without a bias the maximum commutes with the (positive) scale: one FMNMX per element, the multiply happens once at the end
      const float x = RPE ? logit2<HG, RPE>(c, v[e], jk) : __uint_as_float(v[e]);
      m = fmaxf(m, jk < c.len ? x : -INFINITY);
    }
  }
  return RPE ? m : m * c.sl2;
}
template <int HG, bool RPE, bool DROP>
__device__ __forceinline__ float softmax_exp(const RowCtx& c, float m, uint8_t* sP) {
  float l = 0.f;
#pragma unroll 1
  for (int cc = 0; cc < 8; ++cc) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(c.lane_addr + cc * 16, v);
    tmem_ld_wait();
    float pe[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int jk = cc * 16 + e;
      const float x = logit2<HG, RPE>(c, v[e], jk);
      float pv = jk < c.len ? ex2_approx(x - m) : 0.f;
      l += pv;
      if (DROP) pv = attn_drop_keep(c.hq, (uint32_t)jk, c.thr24) ? pv * c.ks : 0.f;
      pe[e] = pv;
    }
#pragma unroll
    for (int w2 = 0; w2 < 2; ++w2) {
      uint4 w;
      w.x = pack2(pe[w2 * 8 + 0], pe[w2 * 8 + 1]); w.y = pack2(pe[w2 * 8 + 2], pe[w2 * 8 + 3]);
      w.z = pack2(pe[w2 * 8 + 4], pe[w2 * 8 + 5]); w.w = pack2(pe[w2 * 8 + 6], pe[w2 * 8 + 7]);
      const int ch = cc * 2 + w2;                                // 16-byte chunk of the 256-byte P row: key chunk ch >> 3
      *(uint4*)(sP + (ch >> 3) * 16384 + c.r * 128 + (((ch & 7) ^ (c.r & 7)) << 4)) = w;
    }
  }
  return l;
}

template <int DHP, int HG>
struct FwdSmem {
  static constexpr int Q_OFF = 0, K_OFF = 16384, V_OFF = 32768;
  static constexpr int P0_OFF = 0;                               // P of head 0 overlays the Q / K tiles (dead once S is in TMEM)
  static constexpr int P1_OFF = 49152;
  static constexpr int MISC_OFF = 49152 + (HG == 2 ? 32768 : 0);
  static constexpr int AFF_OFF = MISC_OFF + 64;                  // float[4][DHP]: q gamma, q beta, k gamma, k beta
  static constexpr int GRID_OFF = AFF_OFF + 4 * DHP * 4;         // int[128][3]
  static constexpr int TAB_OFF = GRID_OFF + 1536;                // float[MAX_TAB][HG]
  static constexpr int PT_OFF = TAB_OFF + MAX_TAB * HG * 4;      // int[128]: destination point of every REAL query row (-1: dead / alias slot)
  static constexpr int TOTAL = PT_OFF + 512 + 1024;
};

template <int DHP, int HG>
__global__ void __launch_bounds__(128 * HG, 2) k_pattn_fwd(const Args a) {
  using L = FwdSmem<DHP, HG>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still a SHARED-space pointer (LDS / STS, not generic LD / ST)
  uint8_t *sQ = smem + L::Q_OFF, *sK = smem + L::K_OFF, *sV = smem + L::V_OFF;
  uint64_t* bars = (uint64_t*)(smem + L::MISC_OFF);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2);
  float* sAff = (float*)(smem + L::AFF_OFF);
  int* sGrid = (int*)(smem + L::GRID_OFF);
  float* sTab = (float*)(smem + L::TAB_OFF);
  int* sPt = (int*)(smem + L::PT_OFF);
  constexpr int NT = 128 * HG;
  using C = Coal<DHP, HG, NT>;

  const int p = blockIdx.x, h0 = blockIdx.y * HG;
  pdl_wait();                                                    // programmatic dependent launch (common.cuh): q / k / v come from the previous kernel
  pdl_launch_dependents();
  const int s0 = a.cu[p], len = a.cu[p + 1] - s0;
  if (len <= 0) return;                                          // uniform per CTA
  const int nh = min(HG, a.H - h0);
  const int t = threadIdx.x, r = t & 127, hf = t >> 7, warp = t >> 5;
  const int Dh = a.Dh;
  PTS(0);
  const bool has_ln = a.qg != nullptr, has_rpe = a.rpe_table != nullptr;
  if (t == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128 * HG);
  const int ntab = has_rpe ? 3 * (2 * a.pos_bnd + 1) : 0;
  for (int i = t; i < ntab * HG; i += 128 * HG) { int row = i / HG, j = i - row * HG; sTab[i] = (j < nh) ? a.rpe_table[row * a.H + h0 + j] : 0.f; }
  const bool live = r < len;
  // ---- gather, coalesced (8 lanes per 128-byte row): all three tensors' loads are issued first, the barrier that publishes the
  // LayerNorm parameters comes after, then normalise + store.  In parallel, thread r < 128 resolves where query row r goes.
  int pts[C::RPT];
  coal_rows<DHP, HG, NT>(a.gather_idx + s0, len, t, pts);
  uint4 rq[C::RPT], rk[C::RPT], rv[C::RPT];
  coal_load<DHP, HG, NT>(a.q, a.ldq, h0 * Dh, pts, nh, Dh, t, rq);
  coal_load<DHP, HG, NT>(a.k, a.ldk, h0 * Dh, pts, nh, Dh, t, rk);
  coal_load<DHP, HG, NT>(a.v, a.ldv, h0 * Dh, pts, nh, Dh, t, rv);
  if (hf == 0) {
    const int pt = live ? a.gather_idx[s0 + r] : 0;
    const bool real_q = live && a.slot_of_point[pt] == s0 + r;   // alias slots of a padded tail are keys only
    sPt[r] = real_q ? pt : -1;
    if (has_rpe && live) { sGrid[r * 3] = a.grid[pt * 3]; sGrid[r * 3 + 1] = a.grid[pt * 3 + 1]; sGrid[r * 3 + 2] = a.grid[pt * 3 + 2]; }
  }
  if (has_ln && t >= NT - DHP) {                                   // LayerNorm parameters -> shared memory, AFTER this warp's row loads are in flight
    const int d = t - (NT - DHP);
    const bool ok = d < Dh;
    sAff[d] = ok ? a.qg[d] : 0.f; sAff[DHP + d] = ok ? a.qb[d] : 0.f;
    sAff[2 * DHP + d] = ok ? a.kg[d] : 0.f; sAff[3 * DHP + d] = ok ? a.kb[d] : 0.f;
  }
  if (a.dbg != nullptr && t == 0 && pts[0] >= -1) a.dbg[(blockIdx.x + gridDim.x * blockIdx.y) * 8 + 6] = clock64();
  if (has_ln) __syncthreads();                                   // the affine parameters are in shared memory
  PTS(7);
  coal_store<DHP, HG, NT>(rq, pts, nh, Dh, has_ln ? sAff : nullptr, sAff + DHP, t, sQ);
  coal_store<DHP, HG, NT>(rk, pts, nh, Dh, has_ln ? sAff + 2 * DHP : nullptr, sAff + 3 * DHP, t, sK);
  coal_store<DHP, HG, NT>(rv, pts, nh, Dh, nullptr, nullptr, t, sV);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  PTS(1);

  // ---- S_j = Q_j K_j^T (accumulator columns [128 j, 128 j + 128))
  if (t == 0) {
    const uint32_t idesc = make_idesc(128, 128, false, false);
    const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
    for (int j = 0; j < nh; ++j) {
#pragma unroll
      for (int kk = 0; kk < DHP / 16; ++kk) {
        const uint32_t off = (uint32_t)(j * DHP * 2 + kk * 32);
        umma_bf16(tmem_base + j * 128, make_desc(qa + off, 16, 1024), make_desc(ka + off, 16, 1024), idesc, kk > 0 ? 1u : 0u);
      }
    }
    umma_commit(&bars[0]);
  }
  mbar_wait_hot(&bars[0], 0);
  tc_fence_after();
  PTS(2);

  // ---- softmax of row r of head hf: logits = S * scale (+ RPE bias), two passes over TMEM in rolled 16-column steps (the unrolled
  // 128-element form did not fit the instruction cache: "no instruction" was the dominant stall), P (bf16, unnormalised) -> shared memory
  const int j = hf;                                              // head slot of this thread
  const bool act = j < nh;
  const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(j * 128);
  float m = -INFINITY, l = 0.f;
  if (act) {
    RowCtx rc;
    rc.lane_addr = lane_addr; rc.len = len; rc.r = r; rc.j = j; rc.sl2 = a.scale * 1.4426950408889634f;
    rc.sGrid = sGrid; rc.sTab = sTab; rc.bnd = a.pos_bnd; rc.num = 2 * a.pos_bnd + 1;
    rc.g0 = has_rpe ? sGrid[r * 3] : 0; rc.g1 = has_rpe ? sGrid[r * 3 + 1] : 0; rc.g2 = has_rpe ? sGrid[r * 3 + 2] : 0;
    rc.ks = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f; rc.thr24 = attn_drop_thr(a.drop_p);
    rc.hq = attn_drop_rowhash(a.seed, a.layer_id * 0x9E37u + (uint32_t)(p * a.H + h0 + j), (uint32_t)r);
    uint8_t* sP = smem + (j == 0 ? L::P0_OFF : L::P1_OFF);
    const bool drop = a.drop_p > 0.f;
    if (has_rpe) {
      m = softmax_max<HG, true>(rc);
      l = drop ? softmax_exp<HG, true, true>(rc, m, sP) : softmax_exp<HG, true, false>(rc, m, sP);
    } else {
      m = softmax_max<HG, false>(rc);
      l = drop ? softmax_exp<HG, false, true>(rc, m, sP) : softmax_exp<HG, false, false>(rc, m, sP);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();            // every thread has drained S from TMEM and P is visible to the async proxy
  tc_fence_after();
  PTS(3);

  // ---- O = P_j V: V is the MN-major B operand straight from its [key][d] rows; N spans both heads' columns (head j keeps its own)
  if (t == 0) {
    const uint32_t idesc = make_idesc(128, HG * DHP, false, true);
    const uint32_t va = smem_u32(sV);
    for (int jj = 0; jj < nh; ++jj) {
      const uint32_t pa = smem_u32(smem + (jj == 0 ? L::P0_OFF : L::P1_OFF));
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        umma_bf16(tmem_base + jj * 128, make_desc(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024), make_desc(va + kk * 2048, 8192, 1024), idesc,
                  kk > 0 ? 1u : 0u);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait_hot(&bars[1], 0);
  tc_fence_after();
  PTS(4);
  // ---- epilogue: row r of head j -> normalised bf16 -> staging tile (over the dead V tile) -> coalesced scatter to the real rows
  uint8_t* sO = sV;
  if (act) {
    constexpr int NLD = (DHP + 31) / 32;
    uint32_t o[NLD * 32];
#pragma unroll
    for (int c = 0; c < NLD; ++c) tmem_ld_32x32b_x32(lane_addr + (uint32_t)(j * DHP + c * 32), o + c * 32);
    tmem_ld_wait();
    if (live) a.lse[(int64_t)(s0 + r) * a.H + h0 + j] = (m + lg2_approx(l)) * 0.6931471805599453f;
    const float inv = 1.f / l;
#pragma unroll
    for (int c = 0; c < DHP / 8; ++c) {
      uint4 w;
      w.x = pack2(__uint_as_float(o[c * 8 + 0]) * inv, __uint_as_float(o[c * 8 + 1]) * inv);
      w.y = pack2(__uint_as_float(o[c * 8 + 2]) * inv, __uint_as_float(o[c * 8 + 3]) * inv);
      w.z = pack2(__uint_as_float(o[c * 8 + 4]) * inv, __uint_as_float(o[c * 8 + 5]) * inv);
      w.w = pack2(__uint_as_float(o[c * 8 + 6]) * inv, __uint_as_float(o[c * 8 + 7]) * inv);
      *(uint4*)(sO + r * 128 + (((j * (DHP / 8) + c) ^ (r & 7)) << 4)) = w;
    }
  }
  tc_fence_before();
  __syncthreads();
  {
    const int ch = t % C::LPR, jj = ch / C::GL, c = ch % C::GL;
    if (jj < nh && c * 8 < Dh) {
#pragma unroll
      for (int i = 0; i < C::RPT; ++i) {
        const int row = t / C::LPR + i * C::RPP;
        const int dst = sPt[row];
        if (dst >= 0) *(uint4*)(a.out + (int64_t)dst * a.ldo + (h0 + jj) * Dh + c * 8) = *(const uint4*)(sO + row * 128 + ((ch ^ (row & 7)) << 4));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  PTS(5);
  if (warp == 1) tmem_dealloc(tmem_base, 128 * HG);
}

template <int DHP, int HG>
static int launch_fwd(const Args& a, int n_patches, cudaStream_t st) {
  constexpr int SMEM = FwdSmem<DHP, HG>::TOTAL;
  static bool attr_set = false;
  if (!attr_set) {
    LOTUS_CUDA_CALL(cudaFuncSetAttribute(k_pattn_fwd<DHP, HG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  launch_pdl(k_pattn_fwd<DHP, HG>, dim3(n_patches, (a.H + HG - 1) / HG), dim3(128 * HG), SMEM, st, a);
  LOTUS_CUDA_OK();
  return 0;
}


// ------------------------------------------------------------------------------------------------ backward
// One CTA = one (patch, head pair), 256 threads PER HEAD: both heads run concurrently (one CTA per SM, 16 warps, all 512 TMEM columns).
// Everything is computed key-major so the probability / score-gradient tiles land in shared memory as K-major A operands:
//   S^T [key x q] = K Q^T          dP^T [key x q] = V dO^T                        (head j: TMEM columns 256 j + [0,128) / [128,256))
//   P^T = exp(S^T * scale + bias - lse[q]) (dropout),   dS^T = P^T * (dP^T - delta[q]) * scale      (thread = (head, key, 64 queries))
//   dV [key x d] = P^T dO          dK [key x d] = dS^T Q          dQ [q x d] = dS K            (columns 256 j + 0 / 64 / 128)
// dO, Q and K are consumed as MN-major B operands from their natural [row][d] tiles (N spans both heads' columns, each head keeps
// its own), dS is the MN-major A view of the dS^T tile.  Two MMA phases, two block barriers.
template <int DHP, int HG>
struct BwdSmem {
  static constexpr int Q_OFF = 0, K_OFF = 16384, V_OFF = 32768, DO_OFF = 49152;
  static constexpr int PT_OFF = 65536, DST_OFF = PT_OFF + HG * 32768;
  static constexpr int MISC_OFF = DST_OFF + HG * 32768;
  static constexpr int LSE_OFF = MISC_OFF + 64;                  // float[HG][128] lse (log2 domain), float[HG][128] delta
  static constexpr int AFF_OFF = LSE_OFF + 2 * HG * 128 * 4;     // float[4][DHP]
  static constexpr int GRID_OFF = AFF_OFF + 4 * DHP * 4;         // int[128][3]
  static constexpr int TAB_OFF = GRID_OFF + 1536;                // float[MAX_TAB][HG] table, then float[MAX_TAB][HG] table gradient
  static constexpr int HQ_OFF = TAB_OFF + 2 * MAX_TAB * HG * 4;  // uint32[HG][128] dropout row hashes (one per head and query)
  static constexpr int ROW_OFF = HQ_OFF + HG * 128 * 4;          // int[128] point of every live row (-1: dead), int[128] 1 = real slot / 0 = alias slot
  static constexpr int TOTAL = ROW_OFF + 1024 + 1024;
  // epilogue overlays (every operand tile is dead once the phase-B MMAs have completed):
  //   V tile   -> dV rows, bf16, K-major swizzled like the operand tiles
  //   P^T tile -> dK-hat rows (gradient w.r.t. the NORMALISED k), fp32 [128][HG * DHP], 16-byte chunks swizzled by (row & 7)
  //   dS^T tile-> dQ-hat rows, same layout
  //   Q tile   -> float[warps][4][HG * DHP] LayerNorm affine-gradient partials
  static_assert((8 * HG) * 4 * HG * DHP * 4 <= 16384, "affine partials must fit the Q tile");
};

// sum over the 32 lanes of a warp of v[0..31]: lane L ends up with the total of element L (31 shuffles instead of 160)
__device__ __forceinline__ float warp_transpose_sum32(float* v, int lane) {
#pragma unroll
  for (int s = 16, n = 32; s >= 1; s >>= 1, n >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = upper ? v[i] : v[i + n / 2];
      const float recv = __shfl_xor_sync(0xffffffffu, send, s);
      v[i] = (upper ? v[i + n / 2] : v[i]) + recv;
    }
  }
  return v[0];
}

// LayerNorm backward of one 8-channel chunk of one (row, head); the GL lanes that hold the head cooperate through shuffles.
// xraw: the raw chunk, dy: gradient w.r.t. the normalised + affine output (in) -> gradient w.r.t. the raw values (out);
// this lane's dgamma / dbeta contributions are accumulated into ag / ab.  All 32 lanes must call it (`on` masks rows / chunks).
template <int GL>
__device__ __forceinline__ void ln_bwd_chunk(const uint4& xraw, bool on, int Dh, const float* __restrict__ sg8, float* dy, float* ag, float* ab) {
  float x[8];
  const __nv_bfloat162* h2 = (const __nv_bfloat162*)&xraw;
#pragma unroll
  for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(h2[e]); x[2 * e] = f.x; x[2 * e + 1] = f.y; }
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) s += on ? x[e] : 0.f;
#pragma unroll
  for (int o = GL >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s / (float)Dh;
  float qv = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) { const float d = on ? x[e] - mu : 0.f; qv += d * d; }
#pragma unroll
  for (int o = GL >> 1; o > 0; o >>= 1) qv += __shfl_xor_sync(0xffffffffu, qv, o);
  const float rs = rsqrtf(qv / (float)Dh + 1e-6f);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float xh = on ? (x[e] - mu) * rs : 0.f;
    const float d = on ? dy[e] : 0.f;
    const float g = d * sg8[e];
    ag[e] += d * xh;
    ab[e] += d;
    s1 += g; s2 += g * xh;
    x[e] = xh; dy[e] = g;
  }
#pragma unroll
  for (int o = GL >> 1; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  s1 /= (float)Dh; s2 /= (float)Dh;
#pragma unroll
  for (int e = 0; e < 8; ++e) dy[e] = on ? rs * (dy[e] - s1 - x[e] * s2) : 0.f;
}

// TMEM columns [col, col + ncols) of this thread's lane -> fp32 staging row (16-byte chunks c0 .., swizzled by the row)
template <int NCOLS>
__device__ __forceinline__ void tmem_to_stage_f32(uint32_t taddr, uint8_t* stage_row, int c0, int r) {
#pragma unroll
  for (int b = 0; b < NCOLS / 16; ++b) {
    uint32_t o[16];
    tmem_ld_32x32b_x16(taddr + b * 16, o);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int ch = c0 + b * 4 + c;
      *(uint4*)(stage_row + (((ch & ~7) | ((ch ^ r) & 7)) << 4)) = make_uint4(o[c * 4], o[c * 4 + 1], o[c * 4 + 2], o[c * 4 + 3]);
    }
  }
}
// one 8-float chunk `ch` of staging row `row` (row pitch W floats)
template <int W>
__device__ __forceinline__ void stage_read8(const uint8_t* stage, int row, int ch, float* v) {
  const uint8_t* rp = stage + row * (W * 4);
  const int c0 = 2 * ch, c1 = 2 * ch + 1;
  const float4 a = *(const float4*)(rp + (((c0 & ~7) | ((c0 ^ row) & 7)) << 4));
  const float4 b = *(const float4*)(rp + (((c1 & ~7) | ((c1 ^ row) & 7)) << 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 w;
  w.x = pack2(v[0], v[1]); w.y = pack2(v[2], v[3]); w.z = pack2(v[4], v[5]); w.w = pack2(v[6], v[7]);
  return w;
}

// score-gradient pass of one thread: key row r of head j against queries q0 .. q0 + 63 (rolled 16-column steps)
struct ColCtx {
  uint32_t s_addr, dp_addr; int len, r, j, q0; float sl2, scale; bool live;
  const float *sLse, *sDelta; const int* sGrid; const float* sTab; float* sDTab; int bnd, num, g0, g1, g2;
  float ks; const uint32_t* sHq; uint32_t kmix, thr24;             // dropout: keep-scale, per-query hashes, this key's index, threshold
};
template <int HG, bool RPE, bool DRPE, bool DROP>
__device__ __forceinline__ void score_grad_pass(const ColCtx& c, uint8_t* sPTrow, uint8_t* sDSTrow) {
#pragma unroll 1
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t sv[16], dv[16];
    tmem_ld_32x32b_x16(c.s_addr + cc * 16, sv);
    tmem_ld_32x32b_x16(c.dp_addr + cc * 16, dv);
    tmem_ld_wait();
    float pd[16], ds[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int qi = c.q0 + cc * 16 + e;
      float x = __uint_as_float(sv[e]) * c.sl2;
      int i0 = 0, i1 = 0, i2 = 0;
      if (RPE) {                                                 // bias index: query minus key, clamped (model.py:400-408)
        i0 = (max(-c.bnd, min(c.bnd, c.sGrid[qi * 3] - c.g0)) + c.bnd) * HG + c.j;
        i1 = (max(-c.bnd, min(c.bnd, c.sGrid[qi * 3 + 1] - c.g1)) + c.bnd + c.num) * HG + c.j;
        i2 = (max(-c.bnd, min(c.bnd, c.sGrid[qi * 3 + 2] - c.g2)) + c.bnd + 2 * c.num) * HG + c.j;
        x += (c.sTab[i0] + c.sTab[i1] + c.sTab[i2]) * 1.4426950408889634f;
      }
      const float pij = c.live ? ex2_approx(x - c.sLse[qi]) : 0.f;   // dead query columns carry lse = +inf
      float pdv = pij, dp = __uint_as_float(dv[e]);
      if (DROP) {
        const bool keep = attn_drop_keep(c.sHq[qi], c.kmix, c.thr24);
        pdv = keep ? pij * c.ks : 0.f;
        dp = keep ? dp * c.ks : 0.f;
      }
      const float dsu = pij * (dp - c.sDelta[qi]);               // gradient of the logit (before the 1/sqrt(d) scale)
      if (DRPE) { if (c.live && qi < c.len && dsu != 0.f) { atomicAdd(&c.sDTab[i0], dsu); atomicAdd(&c.sDTab[i1], dsu); atomicAdd(&c.sDTab[i2], dsu); } }
      pd[e] = pdv;
      ds[e] = dsu * c.scale;
    }
#pragma unroll
    for (int w2 = 0; w2 < 2; ++w2) {
      uint4 w, u;
      w.x = pack2(pd[w2 * 8 + 0], pd[w2 * 8 + 1]); w.y = pack2(pd[w2 * 8 + 2], pd[w2 * 8 + 3]);
      w.z = pack2(pd[w2 * 8 + 4], pd[w2 * 8 + 5]); w.w = pack2(pd[w2 * 8 + 6], pd[w2 * 8 + 7]);
      u.x = pack2(ds[w2 * 8 + 0], ds[w2 * 8 + 1]); u.y = pack2(ds[w2 * 8 + 2], ds[w2 * 8 + 3]);
      u.z = pack2(ds[w2 * 8 + 4], ds[w2 * 8 + 5]); u.w = pack2(ds[w2 * 8 + 6], ds[w2 * 8 + 7]);
      const int off = (((cc * 2 + w2) ^ (c.r & 7)) << 4);
      *(uint4*)(sPTrow + off) = w;
      *(uint4*)(sDSTrow + off) = u;
    }
  }
}

template <int DHP, int HG>
__global__ void __launch_bounds__(256 * HG, 1) k_pattn_bwd(const Args a) {
  using L = BwdSmem<DHP, HG>;
  constexpr int NT = 256 * HG, W = HG * DHP;
  using C = Coal<DHP, HG, NT>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still a SHARED-space pointer (LDS / STS, not generic LD / ST)
  uint8_t *sQ = smem + L::Q_OFF, *sK = smem + L::K_OFF, *sV = smem + L::V_OFF, *sdO = smem + L::DO_OFF;
  uint8_t *sPT = smem + L::PT_OFF, *sDST = smem + L::DST_OFF;
  uint64_t* bars = (uint64_t*)(smem + L::MISC_OFF);
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);
  float* sLse = (float*)(smem + L::LSE_OFF);                     // [HG][128]
  float* sDelta = sLse + HG * 128;                               // [HG][128]
  float* sAff = (float*)(smem + L::AFF_OFF);
  int* sGrid = (int*)(smem + L::GRID_OFF);
  float* sTab = (float*)(smem + L::TAB_OFF);
  float* sDTab = sTab + MAX_TAB * HG;
  float* sRed = (float*)sQ;                                       // epilogue overlay
  uint32_t* sHq = (uint32_t*)(smem + L::HQ_OFF);
  int* sPt = (int*)(smem + L::ROW_OFF);
  int* sReal = sPt + 128;

  const int p = blockIdx.x, h0 = blockIdx.y * HG;
  pdl_wait();
  pdl_launch_dependents();
  const int s0 = a.cu[p], len = a.cu[p + 1] - s0;
  if (len <= 0) return;
  const int nh = min(HG, a.H - h0);
  const int t = threadIdx.x, r = t & 127, hf = (t >> 7) & 1, j = t >> 8, warp = t >> 5, lane = t & 31;
  const int Dh = a.Dh;
  PTS(0);
  const bool has_ln = a.qg != nullptr, has_rpe = a.rpe_table != nullptr;
  if (t == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256 * HG);
  const int ntab = has_rpe ? 3 * (2 * a.pos_bnd + 1) : 0;
  for (int i = t; i < ntab * HG; i += NT) {
    int row = i / HG, jj = i - row * HG;
    sTab[i] = (jj < nh) ? a.rpe_table[row * a.H + h0 + jj] : 0.f;
    sDTab[i] = 0.f;
  }
  const bool live = r < len;
  // ---- gather, coalesced (LPR lanes per 128-byte row): the loads of all five tensors are issued first, the barrier that publishes
  // the LayerNorm parameters follows, then normalise + store.  Alias slots are keys only: their query column is switched off through
  // lse = +inf (P^T column 0, dS^T column 0), which is what a zero dO row did before, without a dependent load in front of dO.
  int pts[C::RPT];
  coal_rows<DHP, HG, NT>(a.gather_idx + s0, len, t, pts);
  uint4 rq[C::RPT], rk[C::RPT], rv[C::RPT], rg[C::RPT], ro[C::RPT];
  coal_load<DHP, HG, NT>(a.q, a.ldq, h0 * Dh, pts, nh, Dh, t, rq);
  coal_load<DHP, HG, NT>(a.k, a.ldk, h0 * Dh, pts, nh, Dh, t, rk);
  coal_load<DHP, HG, NT>(a.v, a.ldv, h0 * Dh, pts, nh, Dh, t, rv);
  coal_load<DHP, HG, NT>(a.dout, a.ldo, h0 * Dh, pts, nh, Dh, t, rg);
  coal_load<DHP, HG, NT>(a.out, a.ldo, h0 * Dh, pts, nh, Dh, t, ro);
  if (t < 128 * HG) {                                            // per-row scalars: thread (head jj, row r)
    const int jj = t >> 7;
    const int pt = live ? a.gather_idx[s0 + r] : 0;
    const bool real = live && a.slot_of_point[pt] == s0 + r;
    if (jj == 0) {
      sPt[r] = live ? pt : -1;
      sReal[r] = real ? 1 : 0;
      if (has_rpe && live) { sGrid[r * 3] = a.grid[pt * 3]; sGrid[r * 3 + 1] = a.grid[pt * 3 + 1]; sGrid[r * 3 + 2] = a.grid[pt * 3 + 2]; }
    }
    // dead and alias query columns: exp2(x - inf) = 0
    sLse[jj * 128 + r] = (real && jj < nh) ? a.lse[(int64_t)(s0 + r) * a.H + h0 + jj] * 1.4426950408889634f : INFINITY;
    sHq[jj * 128 + r] = attn_drop_rowhash(a.seed, a.layer_id * 0x9E37u + (uint32_t)(p * a.H + h0 + jj), (uint32_t)r);
  }
  if (has_ln && t >= NT - DHP) {                            // LayerNorm parameters -> shared memory, AFTER this warp's row loads are in flight
    const int d = t - (NT - DHP);
    const bool ok = d < Dh;
    sAff[d] = ok ? a.qg[d] : 0.f; sAff[DHP + d] = ok ? a.qb[d] : 0.f;
    sAff[2 * DHP + d] = ok ? a.kg[d] : 0.f; sAff[3 * DHP + d] = ok ? a.kb[d] : 0.f;
  }
  if (a.dbg != nullptr && t == 0 && pts[0] >= -1) a.dbg[(blockIdx.x + gridDim.x * blockIdx.y) * 8 + 6] = clock64();
  if (has_ln) __syncthreads();                                   // the affine parameters are in shared memory
  PTS(7);
  coal_store<DHP, HG, NT>(rq, pts, nh, Dh, has_ln ? sAff : nullptr, sAff + DHP, t, sQ);
  coal_store<DHP, HG, NT>(rk, pts, nh, Dh, has_ln ? sAff + 2 * DHP : nullptr, sAff + 3 * DHP, t, sK);
  coal_store<DHP, HG, NT>(rv, pts, nh, Dh, nullptr, nullptr, t, sV);
  coal_store<DHP, HG, NT>(rg, pts, nh, Dh, nullptr, nullptr, t, sdO);
  {                                                              // delta = rowsum(dO * O) per (row, head)
    const int ch = t % C::LPR, jj = ch / C::GL, c = ch % C::GL;
#pragma unroll
    for (int i = 0; i < C::RPT; ++i) {
      const int row = t / C::LPR + i * C::RPP;
      const __nv_bfloat162 *g2 = (const __nv_bfloat162*)&rg[i], *o2 = (const __nv_bfloat162*)&ro[i];
      float dsum = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) { float2 x = __bfloat1622float2(g2[e]), y = __bfloat1622float2(o2[e]); dsum = fmaf(x.x, y.x, fmaf(x.y, y.y, dsum)); }
#pragma unroll
      for (int o = C::GL >> 1; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
      if (c == 0) sDelta[jj * 128 + row] = dsum;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(256 * j);
  const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK), va = smem_u32(sV), da = smem_u32(sdO);
  PTS(1);

  // ---- phase A: S^T_j = K_j Q_j^T (cols 256 j + [0,128)), dP^T_j = V_j dO_j^T (cols 256 j + [128,256))
  if (t == 0) {
    const uint32_t idesc = make_idesc(128, 128, false, false);
    for (int jj = 0; jj < nh; ++jj) {
#pragma unroll
      for (int kk = 0; kk < DHP / 16; ++kk) {
        const uint32_t off = (uint32_t)(jj * DHP * 2 + kk * 32);
        umma_bf16(tmem_base + 256 * jj, make_desc(ka + off, 16, 1024), make_desc(qa + off, 16, 1024), idesc, kk > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int kk = 0; kk < DHP / 16; ++kk) {
        const uint32_t off = (uint32_t)(jj * DHP * 2 + kk * 32);
        umma_bf16(tmem_base + 256 * jj + 128, make_desc(va + off, 16, 1024), make_desc(da + off, 16, 1024), idesc, kk > 0 ? 1u : 0u);
      }
    }
    umma_commit(&bars[0]);
  }
  mbar_wait_hot(&bars[0], 0);
  tc_fence_after();
  PTS(2);

  // ---- thread (head j, key r, queries hf*64 .. +63): P^T and dS^T tiles
  const bool act = j < nh;
  if (act) {
    ColCtx cc;
    cc.s_addr = lane_addr + (uint32_t)(hf * 64); cc.dp_addr = lane_addr + (uint32_t)(128 + hf * 64);
    cc.len = len; cc.r = r; cc.j = j; cc.q0 = hf * 64; cc.sl2 = a.scale * 1.4426950408889634f; cc.scale = a.scale; cc.live = live;
    cc.sLse = sLse + j * 128; cc.sDelta = sDelta + j * 128; cc.sGrid = sGrid; cc.sTab = sTab; cc.sDTab = sDTab;
    cc.bnd = a.pos_bnd; cc.num = 2 * a.pos_bnd + 1;
    cc.g0 = has_rpe ? sGrid[r * 3] : 0; cc.g1 = has_rpe ? sGrid[r * 3 + 1] : 0; cc.g2 = has_rpe ? sGrid[r * 3 + 2] : 0;
    cc.ks = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f; cc.sHq = sHq + j * 128; cc.kmix = (uint32_t)r; cc.thr24 = attn_drop_thr(a.drop_p);
    uint8_t* prow = sPT + j * 32768 + hf * 16384 + r * 128;
    uint8_t* drow = sDST + j * 32768 + hf * 16384 + r * 128;
    const bool drop = a.drop_p > 0.f;
    if (has_rpe) {
      if (a.drpe != nullptr) { if (drop) score_grad_pass<HG, true, true, true>(cc, prow, drow); else score_grad_pass<HG, true, true, false>(cc, prow, drow); }
      else { if (drop) score_grad_pass<HG, true, false, true>(cc, prow, drow); else score_grad_pass<HG, true, false, false>(cc, prow, drow); }
    } else {
      if (drop) score_grad_pass<HG, false, false, true>(cc, prow, drow); else score_grad_pass<HG, false, false, false>(cc, prow, drow);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  PTS(3);
  // ---- phase B (head j in columns 256 j + ...): dV = P^T dO (+0), dK = dS^T Q (+64), dQ = dS K (+128; A = dS^T tile read MN-major)
  if (t == 0) {
    const uint32_t idesc = make_idesc(128, HG * DHP, false, true);
    const uint32_t idesc_mn = make_idesc(128, HG * DHP, true, true);
    for (int jj = 0; jj < nh; ++jj) {
      const uint32_t pa = smem_u32(sPT + jj * 32768), sa = smem_u32(sDST + jj * 32768), d0 = tmem_base + 256 * jj;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        umma_bf16(d0, make_desc(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024), make_desc(da + kk * 2048, 8192, 1024), idesc, kk > 0 ? 1u : 0u);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        umma_bf16(d0 + 64, make_desc(sa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024), make_desc(qa + kk * 2048, 8192, 1024), idesc, kk > 0 ? 1u : 0u);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        umma_bf16(d0 + 128, make_desc(sa + kk * 2048, 16384, 1024), make_desc(ka + kk * 2048, 8192, 1024), idesc_mn, kk > 0 ? 1u : 0u);
    }
    umma_commit(&bars[1]);
  }
  // the raw k / q chunks of the LayerNorm backward (coalesced mapping): loaded while the MMAs run
  uint4 xk[C::RPT], xq[C::RPT];
  if (has_ln) {
    coal_load<DHP, HG, NT>(a.k, a.ldk, h0 * Dh, pts, nh, Dh, t, xk);
    coal_load<DHP, HG, NT>(a.q, a.ldq, h0 * Dh, pts, nh, Dh, t, xq);
  } else {
#pragma unroll
    for (int i = 0; i < C::RPT; ++i) { xk[i] = make_uint4(0, 0, 0, 0); xq[i] = make_uint4(0, 0, 0, 0); }
  }
  mbar_wait_hot(&bars[1], 0);
  tc_fence_after();
  PTS(4);
  // ---- epilogue 1: TMEM -> staging tiles (thread = (head j, row r); the two halves share the dK columns)
  uint8_t *stV = sV, *stK = sPT, *stQ = sDST;
  if (act) {
    constexpr int HALF = DHP >= 32 ? DHP / 2 : DHP;
    if (hf == 0) {
      constexpr int NLD = (DHP + 31) / 32;
      uint32_t o[NLD * 32];
#pragma unroll
      for (int c = 0; c < NLD; ++c) tmem_ld_32x32b_x32(lane_addr + (uint32_t)(j * DHP + c * 32), o + c * 32);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < DHP / 8; ++c) *(uint4*)(stV + r * 128 + (((j * (DHP / 8) + c) ^ (r & 7)) << 4)) = pack8((const float*)o + c * 8);
      tmem_to_stage_f32<HALF>(lane_addr + (uint32_t)(64 + j * DHP), stK + r * (W * 4), j * (DHP / 4), r);
    } else {
      tmem_to_stage_f32<DHP>(lane_addr + (uint32_t)(128 + j * DHP), stQ + r * (W * 4), j * (DHP / 4), r);
      if constexpr (HALF < DHP) tmem_to_stage_f32<DHP - HALF>(lane_addr + (uint32_t)(64 + j * DHP + HALF), stK + r * (W * 4), j * (DHP / 4) + HALF / 4, r);
    }
  }
  tc_fence_before();
  __syncthreads();
  // ---- epilogue 2: coalesced pass (LPR lanes per row): dV copy, LayerNorm backward of dK-hat / dQ-hat, stores to the packed
  // gradients (real slots) or the alias side buffer (alias slots: dk / dv only)
  {
    const int ch = t % C::LPR, jj = ch / C::GL, c = ch % C::GL;
    const bool col_ok = jj < nh && c * 8 < Dh;
    const int hcol = (h0 + jj) * Dh + c * 8, Call = a.H * Dh;
    float agk[8], abk[8], agq[8], abq[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { agk[e] = 0.f; abk[e] = 0.f; agq[e] = 0.f; abq[e] = 0.f; }
#pragma unroll
    for (int i = 0; i < C::RPT; ++i) {
      const int row = t / C::LPR + i * C::RPP;
      const int pt = pts[i];
      const bool real = sReal[row] != 0;
      const bool on_k = pt >= 0 && col_ok, on_q = real && col_ok;
      __nv_bfloat16 *dvp = nullptr, *dkp = nullptr;
      if (on_k) {
        if (real) { dvp = a.dv + (int64_t)pt * a.lddv + hcol; dkp = a.dk + (int64_t)pt * a.lddk + hcol; }
        else { dkp = a.alias_dkv + (int64_t)(s0 + row) * 2 * Call + hcol; dvp = dkp + Call; }
        *(uint4*)dvp = *(const uint4*)(stV + row * 128 + ((ch ^ (row & 7)) << 4));
      }
      float dy[8];
      stage_read8<W>(stK, row, ch, dy);
      if (has_ln) ln_bwd_chunk<C::GL>(xk[i], on_k, Dh, sAff + 2 * DHP + c * 8, dy, agk, abk);
      if (on_k) *(uint4*)dkp = pack8(dy);
      stage_read8<W>(stQ, row, ch, dy);
      if (has_ln) ln_bwd_chunk<C::GL>(xq[i], on_q, Dh, sAff + c * 8, dy, agq, abq);
      if (on_q) *(uint4*)(a.dq + (int64_t)pt * a.lddq + hcol) = pack8(dy);
    }
    if (has_ln && a.dqg != nullptr) {                            // lanes that hold the same chunk: xor-reduce, then one partial row per warp
#pragma unroll
      for (int e = 0; e < 8; ++e) {
#pragma unroll
        for (int o = C::LPR; o < 32; o <<= 1) {
          agk[e] += __shfl_xor_sync(0xffffffffu, agk[e], o); abk[e] += __shfl_xor_sync(0xffffffffu, abk[e], o);
          agq[e] += __shfl_xor_sync(0xffffffffu, agq[e], o); abq[e] += __shfl_xor_sync(0xffffffffu, abq[e], o);
        }
      }
      if (lane < C::LPR) {
        float* dst = sRed + warp * 4 * W + lane * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) { dst[e] = agk[e]; dst[W + e] = abk[e]; dst[2 * W + e] = agq[e]; dst[3 * W + e] = abq[e]; }
      }
    }
  }
  __syncthreads();
  PTS(5);
  // ---- LayerNorm affine gradients: per-warp partials -> one atomic per (parameter, head, channel) per CTA
  if (has_ln && a.dqg != nullptr && t < 4 * W) {
    const int which = t / W, col = t - which * W;                // 0: k gamma, 1: k beta, 2: q gamma, 3: q beta
    const int jj = col / DHP, d = col - jj * DHP;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) v += sRed[(w * 4 + which) * W + col];
    float* dst = which == 0 ? a.dkg : which == 1 ? a.dkb : which == 2 ? a.dqg : a.dqb;
    if (jj < nh && d < Dh && v != 0.f) atomicAdd(&dst[d], v);
  }
  if (a.drpe != nullptr && has_rpe) {
    for (int i = t; i < ntab * HG; i += NT) {
      int row = i / HG, jj = i - row * HG;
      if (jj < nh && sDTab[i] != 0.f) atomicAdd(&a.drpe[row * a.H + h0 + jj], sDTab[i]);
    }
  }
  if (warp == 1) tmem_dealloc(tmem_base, 256 * HG);
}

// fold the alias-slot contributions into the packed gradient: dk[pt] += alias[slot][0], dv[pt] += alias[slot][1]
__global__ void k_alias_fold(const int32_t* __restrict__ gather_idx, const int32_t* __restrict__ slot_of_point, int n_pad, int C,
                             const __nv_bfloat16* __restrict__ alias, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv,
                             int64_t lddk, int64_t lddv) {
  pdl_wait();
  pdl_launch_dependents();
  const int lanes = C / 8;                                       // 16-byte chunks per row
  const int per_block = blockDim.x / lanes;
  const int slot = blockIdx.x * per_block + threadIdx.x / lanes, c = threadIdx.x % lanes;
  if (threadIdx.x >= per_block * lanes || slot >= n_pad) return;
  const int pt = gather_idx[slot];
  if (slot_of_point[pt] == slot) return;                         // a real slot
  const __nv_bfloat16* src = alias + (int64_t)slot * 2 * C;
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    __nv_bfloat16* dst = (w == 0 ? dk + (int64_t)pt * lddk : dv + (int64_t)pt * lddv) + c * 8;
    uint4 x = *(const uint4*)(src + w * C + c * 8), y = *(uint4*)dst;
    const __nv_bfloat162 *x2 = (const __nv_bfloat162*)&x, *y2 = (const __nv_bfloat162*)&y;
    uint4 o;
    uint32_t* o32 = (uint32_t*)&o;
#pragma unroll
    for (int e = 0; e < 4; ++e) { float2 p = __bfloat1622float2(x2[e]), q = __bfloat1622float2(y2[e]); o32[e] = pack2(p.x + q.x, p.y + q.y); }
    *(uint4*)dst = o;
  }
}

template <int DHP, int HG>
static int launch_bwd(const Args& a, int n_patches, int n_pad, cudaStream_t st) {
  constexpr int SMEM = BwdSmem<DHP, HG>::TOTAL;
  static bool attr_set = false;
  if (!attr_set) {
    LOTUS_CUDA_CALL(cudaFuncSetAttribute(k_pattn_bwd<DHP, HG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  launch_pdl(k_pattn_bwd<DHP, HG>, dim3(n_patches, (a.H + HG - 1) / HG), dim3(256 * HG), SMEM, st, a);
  LOTUS_CUDA_OK();
  const int C = a.H * a.Dh, lanes = C / 8;
  const int threads = lanes >= 256 ? lanes : 256, per_block = threads / lanes;
  if (lanes <= 1024) {
    launch_pdl(k_alias_fold, dim3((n_pad + per_block - 1) / per_block), dim3(threads), 0, st, a.gather_idx, a.slot_of_point, n_pad, C, (const __nv_bfloat16*)a.alias_dkv, a.dk, a.dv, a.lddk, a.lddv);
    LOTUS_CUDA_OK();
  }
  return 0;
}

}  // namespace paf

static long long* g_paf_dbg = nullptr;   // dev: phase-timestamp buffer (lotus_attn_debug_buffer)
static bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// Fused patch attention forward (bf16): q / k / v are the RAW projections (e.g. the three column blocks of the packed qkv tensor);
// qn_* / kn_* (fp32 [Dh], all four or none): per-head LayerNorm affine parameters applied to q and k inside the kernel (eps 1e-6);
// rpe_table ([3 * (2 * pos_bnd + 1), H] fp32) + grid ([n, 3] int32): optional relative-position bias added to the scaled logits.
// out [n, ldo] bf16 (scattered by slot_of_point), lse [n_pad, H] fp32.  Patches of at most 128 tokens; Dh % 8 == 0, Dh <= 64.
LOTUS_API int lotus_patch_attn_fused_fwd_bf16(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                                              const float* qn_w, const float* qn_b, const float* kn_w, const float* kn_b,
                                              const int32_t* gather_idx, const int32_t* slot_of_point, const int32_t* cu_seqlens,
                                              int n_patches, int max_len, int H, int Dh, float scale, const float* rpe_table,
                                              const int32_t* grid, int pos_bnd, void* out, int64_t ldo, float* lse, float drop_p,
                                              uint64_t seed, uint32_t layer_id, void* stream_) {
  LOTUS_CHECK_ARG(max_len <= 128, "lotus_patch_attn_fused_fwd_bf16: patches of at most 128 tokens (got %d)", max_len);
  LOTUS_CHECK_ARG(Dh % 8 == 0 && Dh >= 8 && Dh <= 64, "lotus_patch_attn_fused_fwd_bf16: head dim %d must be a multiple of 8 in [8, 64]", Dh);
  LOTUS_CHECK_ARG(al16(q) && al16(k) && al16(v) && al16(out) && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
                  "lotus_patch_attn_fused_fwd_bf16: rows must be 16-byte aligned");
  LOTUS_CHECK_ARG((qn_w != nullptr) == (qn_b != nullptr) && (qn_w != nullptr) == (kn_w != nullptr) && (qn_w != nullptr) == (kn_b != nullptr),
                  "lotus_patch_attn_fused_fwd_bf16: pass all four q/k-norm parameters or none");
  LOTUS_CHECK_ARG(rpe_table == nullptr || (grid != nullptr && 3 * (2 * pos_bnd + 1) <= paf::MAX_TAB), "lotus_patch_attn_fused_fwd_bf16: RPE needs grid and pos_bnd <= 47");
  if (n_patches <= 0) return 0;
  paf::Args a;
  memset(&a, 0, sizeof(a));
  a.q = (const __nv_bfloat16*)q; a.k = (const __nv_bfloat16*)k; a.v = (const __nv_bfloat16*)v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.qg = qn_w; a.qb = qn_b; a.kg = kn_w; a.kb = kn_b;
  a.gather_idx = gather_idx; a.slot_of_point = slot_of_point; a.cu = cu_seqlens; a.H = H; a.Dh = Dh; a.scale = scale;
  a.rpe_table = rpe_table; a.grid = grid; a.pos_bnd = pos_bnd; a.out = (__nv_bfloat16*)out; a.ldo = ldo; a.lse = lse;
  a.drop_p = drop_p; a.seed = seed; a.layer_id = layer_id; a.dbg = g_paf_dbg;
  cudaStream_t st = (cudaStream_t)stream_;
  if (Dh <= 16) return paf::launch_fwd<16, 2>(a, n_patches, st);
  if (Dh <= 32) return paf::launch_fwd<32, 2>(a, n_patches, st);
  return paf::launch_fwd<64, 1>(a, n_patches, st);
}

// Fused patch attention backward (bf16).  Inputs as the forward (raw q / k / v, LayerNorm parameters, RPE) plus the forward's out / lse
// and dout.  Outputs: dq / dk / dv = gradients w.r.t. the RAW q / k / v rows, bf16, written (not accumulated) for every point;
// dqn_* / dkn_* (fp32 [Dh], accumulated): LayerNorm affine gradients; drpe_table (accumulated, optional).
// alias_scratch: device buffer of n_pad * 2 * H * Dh bf16 (the second key-side contribution of alias tail slots; folded in by a small pass).
LOTUS_API int lotus_patch_attn_fused_bwd_bf16(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                                              const float* qn_w, const float* qn_b, const float* kn_w, const float* kn_b,
                                              const int32_t* gather_idx, const int32_t* slot_of_point, const int32_t* cu_seqlens,
                                              int n_patches, int n_pad, int max_len, int H, int Dh, float scale, const float* rpe_table,
                                              const int32_t* grid, int pos_bnd, float* drpe_table, const void* out, const void* dout,
                                              int64_t ldo, const float* lse, void* dq, void* dk, void* dv, int64_t lddq, int64_t lddk,
                                              int64_t lddv, float* dqn_w, float* dqn_b, float* dkn_w, float* dkn_b, void* alias_scratch,
                                              float drop_p, uint64_t seed, uint32_t layer_id, void* stream_) {
  LOTUS_CHECK_ARG(max_len <= 128, "lotus_patch_attn_fused_bwd_bf16: patches of at most 128 tokens (got %d)", max_len);
  LOTUS_CHECK_ARG(Dh % 8 == 0 && Dh >= 8 && Dh <= 64, "lotus_patch_attn_fused_bwd_bf16: head dim %d must be a multiple of 8 in [8, 64]", Dh);
  LOTUS_CHECK_ARG(al16(q) && al16(k) && al16(v) && al16(out) && al16(dout) && al16(dq) && al16(dk) && al16(dv) && al16(alias_scratch) &&
                      ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0,
                  "lotus_patch_attn_fused_bwd_bf16: rows must be 16-byte aligned");
  LOTUS_CHECK_ARG(alias_scratch != nullptr, "lotus_patch_attn_fused_bwd_bf16: alias_scratch is required");
  LOTUS_CHECK_ARG((dqn_w != nullptr) == (dqn_b != nullptr) && (dqn_w != nullptr) == (dkn_w != nullptr) && (dqn_w != nullptr) == (dkn_b != nullptr),
                  "lotus_patch_attn_fused_bwd_bf16: pass all four q/k-norm gradient outputs or none");
  LOTUS_CHECK_ARG(rpe_table == nullptr || (grid != nullptr && 3 * (2 * pos_bnd + 1) <= paf::MAX_TAB), "lotus_patch_attn_fused_bwd_bf16: RPE needs grid and pos_bnd <= 47");
  if (n_patches <= 0) return 0;
  paf::Args a;
  memset(&a, 0, sizeof(a));
  a.q = (const __nv_bfloat16*)q; a.k = (const __nv_bfloat16*)k; a.v = (const __nv_bfloat16*)v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.qg = qn_w; a.qb = qn_b; a.kg = kn_w; a.kb = kn_b;
  a.gather_idx = gather_idx; a.slot_of_point = slot_of_point; a.cu = cu_seqlens; a.H = H; a.Dh = Dh; a.scale = scale;
  a.rpe_table = rpe_table; a.grid = grid; a.pos_bnd = pos_bnd; a.drpe = drpe_table;
  a.out = (__nv_bfloat16*)out; a.dout = (const __nv_bfloat16*)dout; a.ldo = ldo; a.lse = (float*)lse;
  a.dq = (__nv_bfloat16*)dq; a.dk = (__nv_bfloat16*)dk; a.dv = (__nv_bfloat16*)dv; a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
  a.alias_dkv = (__nv_bfloat16*)alias_scratch; a.dqg = dqn_w; a.dqb = dqn_b; a.dkg = dkn_w; a.dkb = dkn_b;
  a.drop_p = drop_p; a.seed = seed; a.layer_id = layer_id; a.dbg = g_paf_dbg;
  cudaStream_t st = (cudaStream_t)stream_;
  if (Dh <= 16) return paf::launch_bwd<16, 2>(a, n_patches, n_pad, st);
  if (Dh <= 32) return paf::launch_bwd<32, 2>(a, n_patches, n_pad, st);
  return paf::launch_bwd<64, 1>(a, n_patches, n_pad, st);
}

// dev tool: device buffer (8 x int64 per CTA) receiving clock64 stamps of the phases of the next fused attention launches; null disables
LOTUS_API int lotus_attn_debug_buffer(void* buf) { g_paf_dbg = (long long*)buf; return 0; }
