// Persistent bf16 linear kernel for sm_100a:  D[M,N] = epilogue( A[M,K] . B[N,K]^T )   (both operands K-major, TMA-fed)
//
// Serves every dense layer of the policy in tensor-core mode (reference call sites: model.py:560-583 MLP, :497-551 qkv / proj,
// :615-625 CPE linear, :713-790 / :817-828 pool / unpool projections, model_ca.py:46-152 q / kv / proj, simple_policy_ptv3.py:113-120
// head first layers) -- forward, data gradient, and the data gradient fused with gelu' -- through lotus_linear_{fwd,dgrad,dgrad_act}_bf16.
//
// Why persistent: these GEMMs have K = 64..256 at the stages that hold the points (1-4 k-blocks per 128-row tile), so a
// one-tile-per-CTA kernel is all prologue (barrier init, TMEM alloc, descriptor fetch) and epilogue; and they are HBM-bound
// (intensity ~ K*N/(K+N) <= 128 FLOP/B), so what matters is keeping loads, MMAs, epilogue math and stores of DIFFERENT tiles in
// flight at once.  One CTA per SM walks tiles t = blockIdx.x, + gridDim.x, ... (n fastest: the CTAs that share an A tile run
// together and hit it in L2):
//
//   warp 0      TMA producer: A / B k-blocks into a STAGES-deep ring (runs ahead across tile boundaries), residual tiles
//               into a 2-deep ring
//   warp 1      tcgen05.mma issuer; accumulators DOUBLE-BUFFERED in TMEM (2 x BN columns): tile i+1's MMAs run while
//               tile i's epilogue drains the other accumulator
//   warps 2-9   epilogue: tcgen05.ld -> bias / GELU / LeakyReLU / gelu' / dropout / residual -> swizzled smem tile -> TMA store
//               (the store of tile i is still in flight while tile i+1's epilogue math runs)
//   warps 10-13 (bias-gradient mode) column sums of the A operand straight from the smem ring: db[k] += sum_m A[m,k], so the
//               data-gradient GEMM of a linear layer also produces its bias gradient and the separate colsum launch disappears
#include "tc_common.cuh"
#include "tc_host.cuh"
#include <stdlib.h>

namespace lin {
using namespace tc;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;
constexpr int MAX_STAGES = 8;
constexpr int EPI_WARP0 = 2, EPI_WARPS = 8, COL_WARP0 = 10, COL_WARPS = 4;
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_LEAKY = 2, ACT_GELU_GRAD = 3 };

struct Params {
  int M, N, K;
  int n_tiles, num_tiles, num_kb, stages;
  int64_t ldo;
  int out_fp32, has_res, has_out2, act;
  float drop_p;
  uint64_t seed;
  uint32_t drop_stream;
  const float* bias;     // [N] or null
  float* colsum;         // [K] fp32, accumulated: column sums of A over all M rows (bias gradient), or null
  int col_warps;         // bias-gradient warps in use (1, 2 or 4): divides `stages`, so a ring stage always belongs to the same warp
  int out_bufs;          // epilogue staging tiles (1 or 2)
};

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int BN>
__global__ void __launch_bounds__(448, 1)
k_lin(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
      const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
      const __grid_constant__ CUtensorMap map_out2, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int B_BYTES = BN * BK * 2;
  const int STAGES = p.stages;
  const int tile_bytes = BM * BN * (p.out_fp32 ? 4 : 2);
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still a SHARED-space pointer (LDS / STS, not generic LD / ST)
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * A_BYTES;
  const int out_bufs = p.out_bufs;                             // epilogue staging tiles (TMA store sources)
  uint8_t* sOut = sB + STAGES * B_BYTES;
  uint8_t* sAux = sOut + out_bufs * tile_bytes;                // residual ring [2] (or the pre-activation tile of out2)
  uint8_t* tail = sAux + (p.has_res ? 2 * tile_bytes : (p.has_out2 ? out_bufs * tile_bytes : 0));
  uint64_t* full = (uint64_t*)tail;
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* tmem_full = empty + MAX_STAGES;                    // [2]
  uint64_t* tmem_empty = tmem_full + 2;                        // [2]
  uint64_t* res_full = tmem_empty + 2;                         // [2]
  uint64_t* res_empty = res_full + 2;                          // [2]
  uint32_t* tmem_slot = (uint32_t*)(res_empty + 2);
  float* sBias = (float*)(tmem_slot + 4);                      // [BN]
  float* sCol = sBias + BN;                                    // [COL_WARPS][num_kb * 64] per-warp partial column sums (bias-gradient mode)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool do_col = p.colsum != nullptr;
  const int num_kb = p.num_kb;

  if (threadIdx.x == 32) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    if (p.has_res) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_res) : "memory");
    if (p.has_out2) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out2) : "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], do_col ? 2 : 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], EPI_WARPS);
      mbar_init(&res_full[b], 1); mbar_init(&res_empty[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);   // two accumulators, power-of-two allocation
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  if (do_col) for (int i = threadIdx.x; i < COL_WARPS * num_kb * BK; i += blockDim.x) sCol[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) touched no global data and may run while
  // the previous kernel of the stream drains; global reads / writes start below.  The dependents of THIS kernel may be scheduled as
  // soon as every CTA is past this point (they wait at their own griddepcontrol.wait until this grid has completed and flushed).
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int box_cols = p.out_fp32 ? 32 : 64;
      // The ring position and phase are CARRIED (STAGES is a run-time value): these role loops are chains of dependent instructions in
      // one thread, and `g % STAGES`, `g / STAGES` plus eight descriptor constructions per k-block cost ~700 cycles -- that, not
      // tcgen05.commit, was the "fixed ~380 ns per k-block" of this pipeline.
      int s = 0, ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int m0 = (t / p.n_tiles) * BM, n0 = (t % p.n_tiles) * BN;
        if (p.has_res) {
          const int rb = it & 1;
          mbar_wait(&res_empty[rb], ((it >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&res_full[rb], (uint32_t)tile_bytes);
          for (int j = 0; j < BN / box_cols; ++j) tma_load_2d(sAux + rb * tile_bytes + j * 16384, &map_res, &res_full[rb], n0 + j * box_cols, m0);
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], A_BYTES + B_BYTES);
          tma_load_2d(sA + s * A_BYTES, &map_a, &full[s], kb * BK, m0);
          tma_load_2d(sB + s * B_BYTES, &map_b, &full[s], kb * BK, n0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      // descriptors are built once; the start-address field (bits 0-13, units of 16 bytes) advances by plain additions: +2 per
      // 16-element K step inside the 128-byte swizzle span, + stage size per ring slot (shared addresses < 2^18: no carry out)
      const uint64_t da0 = make_desc(smem_u32(sA), 16, 1024), db0 = make_desc(smem_u32(sB), 16, 1024);
      int s = 0, ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int ab = it & 1;
        mbar_wait(&tmem_empty[ab], ((it >> 1) & 1) ^ 1);       // the epilogue has drained this accumulator (two tiles ago)
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t da = da0 + (uint64_t)((uint32_t)s * (A_BYTES >> 4)), db = db0 + (uint64_t)((uint32_t)s * ((uint32_t)B_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[s]);                               // stage free once these MMAs have read it
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tmem_full[ab]);                            // accumulator complete
      }
    }
  } else if (warp < EPI_WARP0 + EPI_WARPS) {
    // ===================== epilogue =====================
    const int lane_grp = warp & 3;                              // TMEM lanes [32*lane_grp, +32) belong to this warp
    const int half = (warp - EPI_WARP0) >> 2;                   // warps w and w+4 share a lane group and split the columns
    const int r = lane_grp * 32 + lane;                         // row inside the tile
    const int et = threadIdx.x - EPI_WARP0 * 32;                // 0..255
    const uint32_t sw = (uint32_t)(r & 7);
    const uint32_t thr = drop_thr(p.drop_p);
    const bool has_drop = p.drop_p > 0.f;
    const float keep_scale = has_drop ? drop_scale(thr) : 1.f;
    const float half_ks = 0.5f * keep_scale;
    const uint32_t thr4 = thr * 0x01010101u;
    const bool add_res = p.has_res && p.act != ACT_GELU_GRAD;
    const bool mask_in_float = has_drop && (add_res || p.out_fp32);
    const int box_cols = p.out_fp32 ? 32 : 64;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int m0 = (t / p.n_tiles) * BM, n0 = (t % p.n_tiles) * BN;
      const int ab = it & 1, aph = (it >> 1) & 1;
      const int row = m0 + r;
      // the TMA store that last used this staging tile (two tiles ago; the previous tile with a single buffer) must have
      // finished READING it before it is overwritten -- the most recent store may still be in flight
      if (et == 0) { if (out_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
      uint8_t* sO = sOut + (out_bufs == 2 ? (it & 1) * tile_bytes : 0);
      uint8_t* sOut2 = sAux + (out_bufs == 2 ? (it & 1) * tile_bytes : 0);   // pre-activation copy (only without a residual ring)
      if (et < BN) { int n = n0 + et; sBias[et] = (p.bias && n < p.N) ? p.bias[n] : 0.f; }
      named_bar_sync(1, EPI_WARPS * 32);
      if (lane == 0) { mbar_wait(&tmem_full[ab], aph); if (p.has_res) mbar_wait(&res_full[ab], aph); }
      __syncwarp();
      tc_fence_after();
      const uint8_t* sRes = sAux + ab * tile_bytes;
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(ab * BN + c0), v);
        tmem_ld_wait();
        if (c0 + 32 >= (half + 1) * (BN / 2)) {                 // last read of this accumulator by this warp: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[ab]);
        }
        const int nbase = n0 + c0;
        float f[32];
        const float4* b4 = (const float4*)(sBias + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 bb = b4[j];
          f[4 * j] = __uint_as_float(v[4 * j]) + bb.x; f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bb.y;
          f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bb.z; f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bb.w;
        }
        if (p.has_out2) {                                       // pre-activation copy (bf16), same swizzled layout as the output tile
          uint8_t* box2 = sOut2 + (c0 >> 6) * 16384 + r * 128;
          const int cb2 = (c0 & 63) >> 3;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const float* g = f + 8 * ch;
            uint4 w;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(g[0], g[1]), h1 = __floats2bfloat162_rn(g[2], g[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(g[4], g[5]), h3 = __floats2bfloat162_rn(g[6], g[7]);
            w.x = *(uint32_t*)&h0; w.y = *(uint32_t*)&h1; w.z = *(uint32_t*)&h2; w.w = *(uint32_t*)&h3;
            *(uint4*)(box2 + (((cb2 + ch) ^ sw) << 4)) = w;
          }
        }
        // dropout keep-masks of the 32 columns: one hash per four columns, bytes compared against the threshold in one SIMD op
        uint32_t dm[8];
        if (has_drop) {
          const uint64_t o0 = (uint64_t)((int64_t)row * p.ldo + nbase);
#pragma unroll
          for (int j = 0; j < 8; ++j) dm[j] = __vcmpgeu4(drop_hash4(p.seed, p.drop_stream, o0 + 4 * j), thr4);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) dm[j] = 0xFFFFFFFFu;
        }
        if (p.act == ACT_GELU_GRAD) {                           // f *= gelu'(aux), aux = bf16 tile that came through the residual ring
          const uint8_t* rbox = sRes + (c0 >> 6) * 16384 + r * 128;
          const int cb2 = (c0 & 63) >> 3;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint4 q = *(const uint4*)(rbox + (((cb2 + ch) ^ sw) << 4));
            const __nv_bfloat162* qh = (const __nv_bfloat162*)&q;
#pragma unroll
            for (int t2 = 0; t2 < 4; ++t2) {
              float2 e = __bfloat1622float2(qh[t2]);
              f[8 * ch + 2 * t2] *= gelu_grad_fast(e.x, half_ks); f[8 * ch + 2 * t2 + 1] *= gelu_grad_fast(e.y, half_ks);
            }
          }
        } else if (p.act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = gelu_fast(f[j], half_ks);
        } else if (p.act == ACT_LEAKY) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = (f[j] > 0.f ? f[j] : 0.02f * f[j]) * keep_scale;
        } else if (has_drop) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= keep_scale;
        }
        if (mask_in_float) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) f[4 * j + q] = ((dm[j] >> (8 * q)) & 1u) ? f[4 * j + q] : 0.f;
        }
        if (p.out_fp32) {
          uint8_t* box = sO + (c0 >> 5) * 16384 + r * 128;
          const uint8_t* rbox = sRes + (c0 >> 5) * 16384 + r * 128;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            float4 w = make_float4(f[4 * ch], f[4 * ch + 1], f[4 * ch + 2], f[4 * ch + 3]);
            if (add_res) {
              float4 q = *(const float4*)(rbox + ((ch ^ sw) << 4));
              w.x += q.x; w.y += q.y; w.z += q.z; w.w += q.w;
            }
            *(float4*)(box + ((ch ^ sw) << 4)) = w;
          }
        } else {
          uint8_t* box = sO + (c0 >> 6) * 16384 + r * 128;
          const uint8_t* rbox = sRes + (c0 >> 6) * 16384 + r * 128;
          const int cb = (c0 & 63) >> 3;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            float* g = f + 8 * ch;
            if (add_res) {
              uint4 q = *(const uint4*)(rbox + (((cb + ch) ^ sw) << 4));
              const __nv_bfloat162* qh = (const __nv_bfloat162*)&q;
#pragma unroll
              for (int t2 = 0; t2 < 4; ++t2) { float2 e = __bfloat1622float2(qh[t2]); g[2 * t2] += e.x; g[2 * t2 + 1] += e.y; }
            }
            uint4 w;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(g[0], g[1]), h1 = __floats2bfloat162_rn(g[2], g[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(g[4], g[5]), h3 = __floats2bfloat162_rn(g[6], g[7]);
            w.x = *(uint32_t*)&h0; w.y = *(uint32_t*)&h1; w.z = *(uint32_t*)&h2; w.w = *(uint32_t*)&h3;
            if (has_drop && !mask_in_float) {
              w.x &= __byte_perm(dm[2 * ch], 0, 0x1100); w.y &= __byte_perm(dm[2 * ch], 0, 0x3322);
              w.z &= __byte_perm(dm[2 * ch + 1], 0, 0x1100); w.w &= __byte_perm(dm[2 * ch + 1], 0, 0x3322);
            }
            *(uint4*)(box + (((cb + ch) ^ sw) << 4)) = w;
          }
        }
      }
      fence_proxy_async();                                      // generic-proxy smem writes -> visible to the TMA (async proxy)
      named_bar_sync(2, EPI_WARPS * 32);
      if (et == 0) {
        for (int j = 0; j < BN / box_cols; ++j) {
          if (n0 + j * box_cols >= p.N) break;
          tma_store_2d(&map_out, sO + j * 16384, n0 + j * box_cols, m0);
          if (p.has_out2) tma_store_2d(&map_out2, sOut2 + j * 16384, n0 + j * box_cols, m0);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (p.has_res) mbar_arrive(&res_empty[ab]);             // every warp is past its reads of this residual slot
      }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // global writes complete before the CTA retires
  } else if (do_col) {
    // ===================== bias gradient: column sums of A from the smem ring =====================
    // k-block g of the ring is summed by warp g % CW into that warp's private partial sums (no shared-memory atomics, no
    // block barrier); the warp's lane 0 is the second arrival on `empty[s]` (the MMA's commit is the first).  CW divides
    // STAGES, so a stage is always served by the same warp, which therefore sees every phase of that stage's barrier in
    // order (a warp that skipped uses of a stage could run two phases ahead, where a parity wait aliases).  Each A tile is
    // summed once: by the tile whose n index equals m_tile % n_tiles, which also spreads the duty over the CTAs.
    const int CW = p.col_warps;
    const int cw = warp - COL_WARP0;
    const int c = lane & 7, rg = lane >> 3;                     // 16-byte chunk (8 k-columns) x interleaved row group
    float* myCol = sCol + cw * num_kb * BK;
    int s = -1, ph = 0, turn = -1;                               // ring position / phase / (k-block counter mod CW), carried
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      const int mt = t / p.n_tiles, nt = t % p.n_tiles;
      const bool mine = nt == (mt % p.n_tiles);
      for (int kb = 0; kb < num_kb; ++kb) {
        if (++s == STAGES) { s = 0; ph ^= 1; }
        if (++turn == CW) turn = 0;
        if (turn != cw) continue;
        if (lane == 0) mbar_wait(&full[s], ph);
        __syncwarp();
        if (mine) {
          const uint8_t* a = sA + s * A_BYTES;
          float acc[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 8
          for (int j = 0; j < 32; ++j) {
            const int rr = j * 4 + rg;                          // a quarter-warp reads one whole 128-byte row: conflict-free
            uint4 q = *(const uint4*)(a + rr * 128 + ((c ^ (rr & 7)) << 4));
            const __nv_bfloat162* qh = (const __nv_bfloat162*)&q;
#pragma unroll
            for (int t2 = 0; t2 < 4; ++t2) { float2 e = __bfloat1622float2(qh[t2]); acc[2 * t2] += e.x; acc[2 * t2 + 1] += e.y; }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) { acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8); acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16); }
          if (lane < 8) {
            float4* dst = (float4*)(myCol + kb * BK + c * 8);
            float4 d0 = dst[0], d1 = dst[1];
            d0.x += acc[0]; d0.y += acc[1]; d0.z += acc[2]; d0.w += acc[3]; d1.x += acc[4]; d1.y += acc[5]; d1.z += acc[6]; d1.w += acc[7];
            dst[0] = d0; dst[1] = d1;
          }
          __syncwarp();                                         // every lane is done reading this stage
        }
        if (lane == 0) mbar_arrive(&empty[s]);
      }
    }
    named_bar_sync(3, COL_WARPS * 32);
    for (int i = threadIdx.x - COL_WARP0 * 32; i < p.K; i += COL_WARPS * 32) {
      float v = sCol[i] + sCol[num_kb * BK + i] + sCol[2 * num_kb * BK + i] + sCol[3 * num_kb * BK + i];
      if (v != 0.f) atomicAdd(&p.colsum[i], v);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

static int dev_knob(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

// out / residual / out2: [M, ldo] (ldo = N), bf16 or fp32.  colsum (optional): [K] fp32 accumulator of A's column sums.
static int launch(const void* A, int64_t lda, const void* B, int64_t ldb, Params p, void* out, const void* residual, void* out2,
                  cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0) return 0;
  LOTUS_CHECK_ARG(p.K > 0, "linear_tc: K must be positive");
  const int m_tiles = (p.M + BM - 1) / BM;
  // tile width: 64 for fp32 tiles (64 KB at 128), narrow outputs, and problems too small to fill the SMs with 128-wide tiles
  int BN = 128;
  if (p.out_fp32 || p.N <= 64 || m_tiles * ((p.N + 127) / 128) < kNumSM) BN = 64;
  else if (p.N % 192 == 0 && p.N % 128 != 0 && p.N <= 192 && !residual) BN = 192;                 // qkv of the 64-channel stage: no dead columns
  static int force_bn = -1;
  if (force_bn < 0) force_bn = dev_knob("LOTUS_LIN_BN", 0);
  if (force_bn == 64 || ((force_bn == 128 || force_bn == 192) && !p.out_fp32)) BN = force_bn;
  p.n_tiles = (p.N + BN - 1) / BN;
  p.num_tiles = m_tiles * p.n_tiles;
  p.num_kb = (p.K + BK - 1) / BK;
  p.has_res = residual != nullptr; p.has_out2 = out2 != nullptr;
  if (p.colsum && p.num_kb * BK > 4096) { lotus_set_error("linear_tc: bias-gradient mode supports K <= 4096"); return 2; }
  const int grid = p.num_tiles < kNumSM ? p.num_tiles : kNumSM;
  const int tile_bytes = BM * BN * (p.out_fp32 ? 4 : 2);
  // one staging tile where the epilogue is ALU-bound anyway (GELU + pre-activation copy, gelu' of the fused data gradient):
  // the shared memory goes to ring stages instead
  // Two staging tiles (per output) wherever they leave enough ring stages: with one, the epilogue of tile i + 1 waits until the TMA
  // store of tile i has finished READING the tile -- ~1 us of serialisation per tile on the wide outputs of the MLP GEMMs.
  const size_t stage_bytes = (size_t)A_BYTES + (size_t)BN * BK * 2;
  const size_t budget = 227 * 1024;
  auto fixed_bytes = [&](int bufs) {
    return (size_t)1024 /*alignment slack*/ + (size_t)tile_bytes * bufs +
           (p.has_res ? 2 * (size_t)tile_bytes : (p.has_out2 ? (size_t)bufs * tile_bytes : 0)) + 256 /*barriers*/ + (size_t)BN * 4 +
           (p.colsum ? (size_t)COL_WARPS * p.num_kb * BK * 4 : 0) + 64;
  };
  const int want_stages = p.num_kb >= 4 ? 4 : 2;                 // (bias-gradient mode rounds the ring down to 2 / 4 / 8)
  p.out_bufs = (fixed_bytes(2) + (size_t)want_stages * stage_bytes <= budget) ? 2 : 1;
  const size_t fixed = fixed_bytes(p.out_bufs);
  if (fixed + 2 * stage_bytes > budget) { lotus_set_error("linear_tc: shared-memory budget exceeded"); return 2; }
  int stages = (int)((budget - fixed) / stage_bytes);
  if (p.colsum) stages = stages >= 8 ? 8 : (stages >= 4 ? 4 : 2);
  const int kb_per_cta = p.num_kb * ((p.num_tiles + grid - 1) / grid);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages > kb_per_cta) stages = kb_per_cta < 2 ? 2 : kb_per_cta;
  if (p.colsum) stages = stages >= 8 ? 8 : (stages >= 4 ? 4 : 2);
  static int force_stages = -1;
  if (force_stages < 0) force_stages = dev_knob("LOTUS_LIN_STAGES", 0);
  if (force_stages >= 2 && force_stages <= stages && !p.colsum) stages = force_stages;
  p.stages = stages;
  p.col_warps = stages >= 4 ? 4 : 2;
  const size_t smem = fixed + (size_t)stages * stage_bytes;

  CUtensorMap ma, mb, mo, mr, mo2;
  memset(&mr, 0, sizeof(mr)); memset(&mo2, 0, sizeof(mo2));
  const int box_cols = p.out_fp32 ? 32 : 64;
  int rc = make_map(&ma, A, p.M, p.K, lda, BK, BM);
  if (!rc) rc = make_map(&mb, B, p.N, p.K, ldb, BK, BN);
  if (!rc) rc = make_map(&mo, out, p.M, p.N, p.ldo, box_cols, BM, p.out_fp32 != 0);
  if (!rc && residual) rc = make_map(&mr, residual, p.M, p.N, p.ldo, box_cols, BM, p.out_fp32 != 0);
  if (!rc && out2) rc = make_map(&mo2, out2, p.M, p.N, p.ldo, box_cols, BM, false);
  if (rc) return rc;
  const int threads = p.colsum ? (COL_WARP0 + COL_WARPS) * 32 : COL_WARP0 * 32;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  if (BN == 128) {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(k_lin<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); set = true; }
    cudaLaunchKernelEx(&cfg, k_lin<128>, ma, mb, mo, mr, mo2, p);
  } else if (BN == 192) {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(k_lin<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); set = true; }
    cudaLaunchKernelEx(&cfg, k_lin<192>, ma, mb, mo, mr, mo2, p);
  } else {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(k_lin<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); set = true; }
    cudaLaunchKernelEx(&cfg, k_lin<64>, ma, mb, mo, mr, mo2, p);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { lotus_set_error("linear_tc launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace lin

static lin::Params lin_params(int M, int N, int K, int64_t ldo, int out_fp32) {
  lin::Params p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.ldo = ldo; p.out_fp32 = out_fp32;
  return p;
}

// y[M,N] = [residual +] drop(act(x[M,K] @ w[N,K]^T + bias)) ; x, w bf16 ; y (and residual) bf16 or fp32 ;
// y_preact (bf16 [M,N], optional): receives x @ w^T + bias before the activation (saved for the backward of a fused GELU)
LOTUS_API int lotus_linear_fwd_bf16(const void* x, const void* w, const float* bias, const void* residual, void* y,
                                    void* y_preact, int M, int N, int K, int act, int out_fp32, float drop_p, uint64_t seed,
                                    uint32_t drop_stream, void* stream_) {
  LOTUS_CHECK_ARG(K % 8 == 0, "lotus_linear_fwd_bf16: K=%d must be a multiple of 8 (16-byte TMA rows)", K);
  LOTUS_CHECK_ARG(y_preact == nullptr || (!out_fp32 && residual == nullptr),
                  "lotus_linear_fwd_bf16: y_preact needs a bf16 output and no residual");
  LOTUS_CHECK_ARG(N % (out_fp32 ? 4 : 8) == 0, "lotus_linear_fwd_bf16: N=%d must keep output rows 16-byte aligned", N);
  lin::Params p = lin_params(M, N, K, N, out_fp32);
  p.bias = bias; p.act = act; p.drop_p = drop_p; p.seed = seed; p.drop_stream = drop_stream;
  return lin::launch(x, K, w, K, p, y, residual, y_preact, (cudaStream_t)stream_);
}

// dx[M,K] = [residual +] dy[M,N] @ w[N,K]   using the pre-transposed weight wt[K,N] (bf16);
// db (optional, [N] fp32): += column sums of dy (the layer's bias gradient), produced by the same kernel
LOTUS_API int lotus_linear_dgrad_bias_bf16(const void* dy, const void* wt, const void* residual, void* dx, float* db, int M, int N,
                                           int K, int out_fp32, void* stream_) {
  LOTUS_CHECK_ARG(N % 8 == 0, "lotus_linear_dgrad_bf16: N=%d must be a multiple of 8", N);
  LOTUS_CHECK_ARG(K % (out_fp32 ? 4 : 8) == 0, "lotus_linear_dgrad_bf16: K=%d must keep output rows 16-byte aligned", K);
  lin::Params p = lin_params(M, K, N, K, out_fp32);
  p.colsum = db;
  return lin::launch(dy, N, wt, N, p, dx, residual, nullptr, (cudaStream_t)stream_);
}
LOTUS_API int lotus_linear_dgrad_bf16(const void* dy, const void* wt, const void* residual, void* dx, int M, int N, int K,
                                      int out_fp32, void* stream_) {
  return lotus_linear_dgrad_bias_bf16(dy, wt, residual, dx, nullptr, M, N, K, out_fp32, stream_);
}

// dx[M,K] = drop(dy[M,N] @ w[N,K] * act'(preact[M,K])) : the backward of `y = drop(act(preact))` fused into the data-gradient GEMM
// of the NEXT layer (act 1 = GELU; preact and dx bf16; the dropout mask is the forward's: same seed / stream / element index);
// db (optional, [N] fp32): += column sums of dy
LOTUS_API int lotus_linear_dgrad_act_bias_bf16(const void* dy, const void* wt, const void* preact, void* dx, float* db, int M, int N,
                                               int K, int act, float drop_p, uint64_t seed, uint32_t drop_stream, void* stream_) {
  LOTUS_CHECK_ARG(N % 8 == 0 && K % 8 == 0, "lotus_linear_dgrad_act_bf16: N=%d, K=%d must be multiples of 8", N, K);
  LOTUS_CHECK_ARG(act == 1 && preact != nullptr, "lotus_linear_dgrad_act_bf16: act must be 1 (GELU) with a pre-activation tensor");
  lin::Params p = lin_params(M, K, N, K, 0);
  p.act = lin::ACT_GELU_GRAD; p.drop_p = drop_p; p.seed = seed; p.drop_stream = drop_stream; p.colsum = db;
  return lin::launch(dy, N, wt, N, p, dx, preact, nullptr, (cudaStream_t)stream_);
}
LOTUS_API int lotus_linear_dgrad_act_bf16(const void* dy, const void* wt, const void* preact, void* dx, int M, int N, int K, int act,
                                          float drop_p, uint64_t seed, uint32_t drop_stream, void* stream_) {
  return lotus_linear_dgrad_act_bias_bf16(dy, wt, preact, dx, nullptr, M, N, K, act, drop_p, seed, drop_stream, stream_);
}
