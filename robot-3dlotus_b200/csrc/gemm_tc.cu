// bf16 tensor-core GEMM family for sm_100a: tcgen05.mma with TMEM accumulators, TMA-fed operands
// (cp.async.bulk.tensor, SWIZZLE_128B), gathered operands staged by cp.async for the sparse conv.
//
//   D[M,N] (+)= sum_k A[m,k] * B[n,k]
//
// One 128 x BN output tile per CTA, BK = 64, 4-stage mbarrier pipeline, warp roles:
//   warp 0 ........ TMA producer (one elected lane)
//   warp 1 ........ TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2-5 ..... (gather modes) cp.async producers of the gathered operand; then the epilogue:
//                   tcgen05.ld of the accumulator -> bias / GELU / LeakyReLU / dropout / residual -> global
//
// Operand modes (reference contractions they serve are listed in include/lotus.h):
//   OP_K_TMA ....... operand stored [rows][K] (K contiguous)         -> K-major smem tile via TMA
//   OP_K_GATHER .... rows gathered through an index table (-1 -> 0)  -> K-major smem tile via cp.async
//   OP_MN_TMA ...... operand stored [K][rows] (rows contiguous)      -> MN-major smem tile via TMA
//   OP_MN_GATHER ... K-rows gathered through an index table          -> MN-major smem tile via cp.async
#include "tc_common.cuh"
#include "tc_host.cuh"
#include <stdlib.h>

namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;            // 64 bf16 = 128 B = one SWIZZLE_128B span
constexpr int MAX_STAGES = 6;
constexpr int NUM_THREADS = 320;  // 10 warps: TMA, MMA, 8 epilogue (warps 2-5 double as cp.async gather producers)
constexpr int EPI_THREADS = 128;  // gather-producer threads (warps 2-5)

enum { OP_K_TMA = 0, OP_K_GATHER = 1, OP_MN_TMA = 2, OP_MN_GATHER = 3, OP_K_TMAGATHER = 4 };   // 4: rows gathered by the TMA unit (tile::gather4)
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_LEAKY = 2, ACT_GELU_GRAD = 3 };   // 3: multiply by gelu'(aux) where aux rides in `residual`

struct GatherSpec {
  const __nv_bfloat16* base;   // rows of `row_elems` bf16
  const int32_t* idx;          // K-major gather: idx[(kb / kb_per_tap) * idx_stride + m]; MN gather: idx[k]
  int64_t idx_stride;
  int row_elems;               // elements per source row (Cin)
  int kb_per_tap;              // k-blocks per kernel tap (Cin / 64)
  int mirror_taps;             // K-major gather: use tap (ntaps-1-d) (conv dgrad)
  int ntaps;
  int idx_in_smem;             // K-major gather: the tile's index rows are staged in shared memory (set by launch())
  int src_rows;                // rows of the gathered tensor (TMA gather: indices >= src_rows read as zeros)
};

struct Params {
  int M, N, K;                 // logical GEMM extents (K = reduction)
  int a_mode, b_mode;
  int kb_begin_stride;         // k-blocks per split (split-K over blockIdx.z)
  unsigned long long* dbg_ts;  // dev knob: per-CTA timestamps (8 slots), set via lotus_tc_debug_buffer
  int debug;                   // dev knob (LOTUS_TC_DEBUG): 1 skip epilogue math+store, 2 skip mainloop, 4 no store wait
  int pair_commit;             // one `empty` barrier (and one tcgen05.commit) per PAIR of stages
  int async_arrive;            // gather producers signal `full` with cp.async.mbarrier.arrive.noinc (no wait_group lag)
  int ldg_gather;              // K-major gather through registers: 8 lanes load one whole 128-byte row (coalesced), then STS
  int stages;                  // smem pipeline depth (2..4): shallow pipelines leave room for several CTAs per SM
  // pair-list mode (sparse-conv weight gradient): blockIdx.z = tap * splits + split, K read from the device
  const int32_t* k_dev;        // [taps] reduction extent per tap, or null
  int splits;
  int64_t tap_idx_stride;      // gather index arrays advance by this per tap
  int64_t tap_out_stride;      // output pointer advances by this (elements) per tap
  GatherSpec ga, gb;
  // epilogue
  void* out;                   // [M, ldo] bf16 or fp32
  void* out2;                  // optional second output (pre-activation), same dtype/ld as out
  int64_t ldo;
  int out_fp32;
  int atomic;                  // fp32 atomicAdd accumulate (split-K)
  const float* bias;           // [N] or null
  float* rowsum_a;             // [M] fp32, accumulated: sum over the reduction of A (weight-gradient GEMM: the BIAS gradient, db = dy^T 1), or null
  const void* residual;        // [M, ldo] same dtype as out, or null
  int act;
  float drop_p;
  uint64_t seed;
  uint32_t drop_stream;
};

// ------------------------------------------------------------------------------------------------ kernel
// smem per stage: A tile 128 x 64 bf16 = 16 KB, B tile BN x 64 bf16.
// K-major tile:  row r (m or n) at r*128 B, 16-B chunk c at (c ^ (r & 7)) * 16          (SBO = 1024)
// MN-major tile: block j = rows/64, k-row r at j*8192 + r*128 B, chunk c swizzled the same way
//                (LBO = 8192 between 64-wide MN blocks, SBO = 1024 between 8-k-row groups)
// stage one BK x WIDTH MN-major tile whose k-rows are gathered through idx[] (rows >= kext are zero)
template <int WIDTH>
__device__ __forceinline__ void stage_mn_gather(uint8_t* tile, const GatherSpec& g, int src_row, int mn0, int mn_extent, int t) {
  int r = t >> 1, half = t & 1;
  const __nv_bfloat16* src = g.base + (int64_t)max(src_row, 0) * g.row_elems + mn0;
  constexpr int CH = WIDTH / 8;                 // 16-B chunks per row
#pragma unroll
  for (int cc = 0; cc < CH / 2; ++cc) {
    int c = half * (CH / 2) + cc;
    int blk = c >> 3, cw = c & 7;
    bool ok = src_row >= 0 && (mn0 + c * 8) < mn_extent;
    cp_async_16(tile + blk * 8192 + r * 128 + ((cw ^ (r & 7)) << 4), src + c * 8, ok);
  }
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 3)
k_gemm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
          const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
          const __grid_constant__ CUtensorMap map_out2, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  const int STAGES = p.stages;
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still a SHARED-space pointer (LDS / STS, not generic LD / ST)
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = (uint64_t*)(sB + STAGES * B_BYTES);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* tmem_full = empty + MAX_STAGES;
  uint64_t* res_full = tmem_full + 1;
  uint32_t* tmem_slot = (uint32_t*)(res_full + 1);
  float* sBias = (float*)(tmem_slot + 4);                                                 // [BN] bias slice of this tile
  uint8_t* sRes = (uint8_t*)(((uintptr_t)(sBias + BN) + 1023) & ~(uintptr_t)1023);        // residual tile (only when p.residual)
  // K-major gather (sparse conv): this tile's rows of the neighbour table, [ntaps][128] int32, staged once by the producer
  // threads -- the per-k-block index lookup was a dependent global load in front of every cp.async group
  int32_t* sIdx = (int32_t*)(sRes + (p.residual ? (size_t)BM * BN * (p.out_fp32 ? 4 : 2) : 0));
  // Bias gradient inside the weight-gradient GEMM: db[n] = sum_m dy[m, n] is the same contraction with x replaced by ones, so one extra
  // MMA per K step multiplies the A tile with a 16-column tile of ones (2 KB, here) into 16 spare TMEM columns; every column holds db.
  // (The column sums used to ride in the data-gradient GEMM on CUDA cores, where they cost as much as the GEMM itself.)
  uint8_t* sOnes = sRes;                                       // only with p.rowsum_a (then there is no residual and no index tile)
  const uint32_t tmem_cols = p.rowsum_a ? 2 * BN : BN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (threadIdx.x == 0) TS(0);
  if (threadIdx.x == 32) {   // descriptor fetches overlap the barrier / TMEM set-up (short tiles are prologue-latency bound)
    if (p.a_mode == OP_K_TMA || p.a_mode == OP_MN_TMA || p.a_mode == OP_K_TMAGATHER) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    if (p.b_mode == OP_K_TMA || p.b_mode == OP_MN_TMA) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    if (p.residual) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_res) : "memory");
    if (p.out2) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out2) : "memory");
  }
  const bool a_gather = (p.a_mode == OP_K_GATHER || p.a_mode == OP_MN_GATHER);
  const bool b_gather = (p.b_mode == OP_MN_GATHER);
  const bool a_mn = (p.a_mode == OP_MN_TMA || p.a_mode == OP_MN_GATHER), b_mn = (p.b_mode >= OP_MN_TMA);
  const bool a_tmag = p.a_mode == OP_K_TMAGATHER;
  const int PAIR = p.pair_commit;

  if (threadIdx.x == 0) {
    // TMA producer + the gather producers: every thread's cp.async completion (async mode) or one elected lane per warp
    const bool warp_arrive = !p.async_arrive;
    uint32_t full_count = 1 + ((a_gather || b_gather) ? (warp_arrive ? EPI_THREADS / 32 : EPI_THREADS) : 0);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], full_count); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(res_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);   // BN in {64,128}: power of two >= 32
  if (p.rowsum_a && threadIdx.x >= 64 && threadIdx.x < 64 + 128) {
    ((uint4*)sOnes)[threadIdx.x - 64] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);   // bf16 1.0
    fence_proxy_async();                               // generic-proxy stores -> visible to the MMA's async proxy
  }
  // programmatic dependent launch: nothing above read global data; from here on the previous kernel's results are needed
  pdl_wait();
  pdl_launch_dependents();
  // reduction range of this CTA
  int tap = 0, Kext = p.K, kb0, kb1;
  if (p.k_dev) {
    tap = blockIdx.z / p.splits;
    int split = blockIdx.z - tap * p.splits;
    Kext = p.k_dev[tap];
    int nkb = (Kext + BK - 1) / BK;
    int per = (nkb + p.splits - 1) / p.splits;
    kb0 = split * per;
    kb1 = min(nkb, kb0 + per);
  } else {
    int nkb = (p.K + BK - 1) / BK;
    kb0 = blockIdx.z * p.kb_begin_stride;
    kb1 = min(nkb, kb0 + p.kb_begin_stride);
  }
  const int num_kb = max(0, kb1 - kb0);
  if (threadIdx.x >= 64 && threadIdx.x - 64 < BN) {   // stage the bias slice once (one coalesced load instead of per-element loads)
    int n = n0 + threadIdx.x - 64;
    sBias[threadIdx.x - 64] = (p.bias && n < p.N) ? p.bias[n] : 0.f;
  }
  if ((p.a_mode == OP_K_GATHER || a_tmag) && p.ga.idx_in_smem && threadIdx.x >= 64 && threadIdx.x < 64 + EPI_THREADS) {
    const int t = threadIdx.x - 64, m = m0 + t;
    for (int tp0 = 0; tp0 < p.ga.ntaps; tp0 += 9) {          // nine independent loads in flight, then the stores
      int v[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        int tp = tp0 + j, src_tp = p.ga.mirror_taps ? p.ga.ntaps - 1 - tp : tp;
        v[j] = (tp < p.ga.ntaps && m < p.M) ? p.ga.idx[(int64_t)src_tp * p.ga.idx_stride + m] : -1;
        if (a_tmag && v[j] < 0) v[j] = p.ga.src_rows;           // absent neighbour -> an out-of-range row: the TMA zero-fills it
      }
#pragma unroll
      for (int j = 0; j < 9; ++j) if (tp0 + j < p.ga.ntaps) sIdx[(tp0 + j) * EPI_THREADS + t] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TS(1);

  if (num_kb > 0 && !(p.debug & 2)) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (a_tmag) {
        // ---- A rows gathered by the TMA unit: lane j issues ONE tile::gather4 for rows 4j..4j+3 of the tile per k-block
        // (the cp.async form of this gather is bound by the LSU: one 16-byte request per cycle per SM, zero-fills included)
        if (lane == 0 && p.residual) {
          const int box_cols = p.out_fp32 ? 32 : 64;
          mbar_arrive_expect_tx(res_full, (uint32_t)(BM * BN * (p.out_fp32 ? 4 : 2)));
          for (int j = 0; j < BN / box_cols; ++j) tma_load_2d(sRes + j * 16384, &map_res, res_full, n0 + j * box_cols, m0);
        }
        for (int i = 0; i < num_kb; ++i) {
          int s = i % STAGES, ph = (i / STAGES) & 1;
          int kb = kb0 + i;
          int tp = kb / p.ga.kb_per_tap, c0 = (kb - tp * p.ga.kb_per_tap) * BK;
          int4 rows = *reinterpret_cast<const int4*>(sIdx + tp * EPI_THREADS + lane * 4);
          if (lane == 0) {
            mbar_wait_hot(&empty[PAIR ? (s >> 1) : s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], A_BYTES + B_BYTES);
            tma_load_2d(sB + s * B_BYTES, &map_b, &full[s], kb * BK, n0);
          }
          __syncwarp();
          tma_gather4_2d(sA + s * A_BYTES + lane * 512, &map_a, &full[s], c0, rows.x, rows.y, rows.z, rows.w);
        }
      } else if (lane == 0) {
        if (p.residual) {   // residual tile [128 x BN] (same dtype as the output), boxes of 128-byte rows
          const int box_cols = p.out_fp32 ? 32 : 64;
          mbar_arrive_expect_tx(res_full, (uint32_t)(BM * BN * (p.out_fp32 ? 4 : 2)));
          for (int j = 0; j < BN / box_cols; ++j) tma_load_2d(sRes + j * 16384, &map_res, res_full, n0 + j * box_cols, m0);
        }
        uint32_t tx = (a_gather ? 0 : A_BYTES) + (b_gather ? 0 : B_BYTES);
        if (p.debug & 128) tx -= B_BYTES;                  // dev knobs: no B / no A loads (timing experiments only)
        if ((p.debug & 512) && !a_gather) tx -= A_BYTES;
        // (ring position and phase are carried, not recomputed: STAGES is a run-time value, and an integer division in a
        // single-thread loop is ~40 dependent instructions -- the k-block rate of these role loops is set by instruction latency)
        int s = 0, ph = 0, k0 = kb0 * BK;
        for (int i = 0; i < num_kb; ++i, k0 += BK, ph ^= (s + 1 == STAGES), s = (s + 1 == STAGES) ? 0 : s + 1) {
          mbar_wait_hot(&empty[PAIR ? (s >> 1) : s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], tx);
          if (!a_gather && !(p.debug & 512)) {
            if (!a_mn) tma_load_2d(sA + s * A_BYTES, &map_a, &full[s], k0, m0);
            else { tma_load_2d(sA + s * A_BYTES, &map_a, &full[s], m0, k0); tma_load_2d(sA + s * A_BYTES + 8192, &map_a, &full[s], m0 + 64, k0); }
          }
          if (!b_gather && !(p.debug & 128)) {
            if (!b_mn) tma_load_2d(sB + s * B_BYTES, &map_b, &full[s], k0, n0);
            else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + s * B_BYTES + j * 8192, &map_b, &full[s], n0 + j * 64, k0);
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      if (lane == 0) {
        // (measured and rejected on B200, all within +-8 % of ~380 ns per k-block for a TMA-fed 4-stage pipeline, even with the
        // loads and the MMAs knocked out: 6 stages, warp-uniform role loops, one commit per two stages, short try_wait hints)
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
                         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        // Descriptors: built once; the start-address field (bits 0-13, units of 16 bytes) is advanced by plain additions
        // (shared-memory addresses stay below 2^18, so no carry leaves the field).  K-major: +32 B per 16-element K step inside
        // the 128-B swizzle span; MN-major: +2 k-groups of 1024 B.  The loop carries ring position and phase: as a chain of
        // dependent instructions in ONE thread, the old form (i % STAGES, i / STAGES, eight make_desc per k-block) was ~700 cycles per
        // k-block, which -- not tcgen05.commit -- was the "fixed ~380 ns per k-block" every variant of this pipeline ran at.
        const uint64_t da0 = a_mn ? make_desc(smem_u32(sA), 8192, 1024) : make_desc(smem_u32(sA), 16, 1024);
        const uint64_t db0 = b_mn ? make_desc(smem_u32(sB), 8192, 1024) : make_desc(smem_u32(sB), 16, 1024);
        const uint32_t a_kstep = a_mn ? (2048u >> 4) : (32u >> 4), b_kstep = b_mn ? (2048u >> 4) : (32u >> 4);
        const bool gen_fence = (a_gather || b_gather) && p.async_arrive;
        // A x ones: N = 16, B read K-major from the all-ones tile (any K step reads ones, so the descriptor does not advance)
        const uint32_t idesc_ones = (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint64_t d_ones = make_desc(smem_u32(sOnes), 16, 1024);
        int s = 0, ph = 0;
        for (int i = 0; i < num_kb; ++i) {
          mbar_wait_hot(&full[s], ph);
          if (gen_fence) fence_proxy_async();   // cp.async (generic proxy) data -> the MMA's async proxy
          if (!(p.debug & 1024)) tc_fence_after();
          if (i == 0) TS(2);
          const uint64_t da = da0 + (uint64_t)((uint32_t)s * (A_BYTES >> 4)), db = db0 + (uint64_t)((uint32_t)s * (B_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (!(p.debug & 256)) umma_bf16(tmem_base, da + (uint64_t)(k * a_kstep), db + (uint64_t)(k * b_kstep), idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          if (p.rowsum_a) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_base + BN, da + (uint64_t)(k * a_kstep), d_ones, idesc_ones, (i > 0 || k > 0) ? 1u : 0u);
          }
          // frees the smem stage(s) when these MMAs retire; pair mode: ONE commit per two stages (tcgen05.commit arrivals
          // retire at a fixed ~380 ns each -- the whole pipeline advanced one k-block per commit whatever it loaded or multiplied)
          if (!PAIR || (s & 1) || i == num_kb - 1) umma_commit(&empty[PAIR ? (s >> 1) : s]);
          if (i == num_kb - 1) { umma_commit(tmem_full); TS(3); }  // accumulator complete
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else if ((a_gather || b_gather) && (warp < 6 || p.async_arrive)) {
      // ===================== cp.async gather producers =====================
      // async-arrive mode: TWO producer groups (warps 2-5 take the even k-blocks, warps 6-9 the odd ones).  A producer
      // thread's own loop iteration (poll `empty`, index, eight copies, arrive) takes ~600 ns whatever the pipeline depth --
      // with copies and B loads knocked out the loop ran at the same rate -- so the k-blocks are dealt to two sets of threads.
      const int t = (threadIdx.x - 64) & (EPI_THREADS - 1);   // 0..127: row of the tile this thread stages
      const int grp = p.async_arrive ? (threadIdx.x - 64) >> 7 : 0, ngrp = p.async_arrive ? 2 : 1;
      const int LAG = STAGES >= 5 ? STAGES - 2 : (STAGES >= 3 ? 2 : 1);   // cp.async groups in flight per thread before signalling
      const int32_t* a_idx = p.ga.idx + (int64_t)tap * p.tap_idx_stride;
      const int32_t* b_idx = p.gb.idx + (int64_t)tap * p.tap_idx_stride;
      // MN-major gathers (pair-list weight gradient): the source row of each k-row comes from an index table; its load is
      // issued two k-blocks ahead so the cp.async never waits on a dependent global load (-25 % on that kernel).
      auto a_row_of = [&](int i) -> int {
        if (i >= num_kb) return -1;
        int kb = kb0 + i;
        if (p.a_mode == OP_MN_GATHER) { int kpt = kb * BK + (t >> 1); return (kpt < Kext) ? a_idx[kpt] : -1; }
        return -1;
      };
      auto b_row_of = [&](int i) -> int {
        if (!b_gather || i >= num_kb) return -1;
        int kpt = (kb0 + i) * BK + (t >> 1);
        return (kpt < Kext) ? b_idx[kpt] : -1;
      };
      const bool mn_rows = p.a_mode == OP_MN_GATHER || b_gather;    // only the MN-major gathers look rows up per k-block
      // K-major gather, two forms.  cp.async (LDGSTS): asynchronous, but the LSU takes ONE 16-byte request per cycle, absent rows
      // (zero fills) included -- 1024 cycles per k-block per SM, shared by the resident CTAs.  LDG.128 with eight lanes per row +
      // STS.128: a quarter of the requests and none for absent rows, but the loads are waited for.  Measured (tools/conv_bench.py):
      // the sparse 96k-row stage (23 % of the (row, tap) pairs present, 5 waves of CTAs) is 5-8 % faster on the first, every deeper
      // stage (36-46 % present, one wave) 8-15 % faster on the second; launch() picks by the row count (p.ldg_gather).
      const bool ldg_path = p.a_mode == OP_K_GATHER && p.ga.idx_in_smem && p.async_arrive && p.ldg_gather;
      int a_r0 = -1, a_r1 = -1, b_r0 = -1, b_r1 = -1;
      if (mn_rows) { a_r0 = a_row_of(grp); a_r1 = a_row_of(grp + ngrp); b_r0 = b_row_of(grp); b_r1 = b_row_of(grp + ngrp); }
      // ring position / phase / (tap, channel block) of k-block i are carried along (no run-time divisions in the loop: eight
      // producer warps x ~330 instructions per k-block made these kernels ISSUE-bound at ~0.5 us per k-block per SM)
      int s = grp % STAGES, ph = (grp / STAGES) & 1;
      int tp_c = 0, cb_c = 0;                                       // K-major gather: tap and 64-channel block of k-block kb0 + i
      if (p.a_mode == OP_K_GATHER) { tp_c = (kb0 + grp) / p.ga.kb_per_tap; cb_c = (kb0 + grp) - tp_c * p.ga.kb_per_tap; }
      int s_lag = 0;                                                // commit/wait_group mode: stage of k-block i - LAG
      for (int i = grp; i < num_kb + (p.async_arrive ? 0 : LAG); i += ngrp) {
        if (i < num_kb) {
          int a_row = -1, b_row = -1;
          if (mn_rows) {
            a_row = a_r0; b_row = b_r0;
            a_r0 = a_r1; b_r0 = b_r1;
            a_r1 = a_row_of(i + 2 * ngrp); b_r1 = b_row_of(i + 2 * ngrp);
          }
          // one lane per warp polls the barrier word (fewer spinning threads competing for issue slots)
          if (lane == 0) mbar_wait_hot(&empty[PAIR ? (s >> 1) : s], ph ^ 1);
          __syncwarp();
          if (ldg_path) {
            // Coalesced gather: EIGHT LANES read one whole 128-byte source row (lane = 16-byte chunk), four rows per warp
            // instruction, through registers.  With one row per lane (the cp.async form below) every warp instruction touches 32
            // different lines = 32 L1TEX wavefronts, and the gather -- not the tensor pipe -- sets the k-block rate.
            const int tp = tp_c, c0 = cb_c * BK;
            const int ch = t & 7, r0 = (t >> 5) * 32 + ((t & 31) >> 3);
            uint4 v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int row = r0 + q * 4;
              const int src_row = sIdx[tp * EPI_THREADS + row];
              v[q] = make_uint4(0, 0, 0, 0);
              if (src_row >= 0) v[q] = *(const uint4*)(p.ga.base + (int64_t)src_row * p.ga.row_elems + c0 + ch * 8);
            }
            uint8_t* dstt = sA + s * A_BYTES;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int row = r0 + q * 4;
              *(uint4*)(dstt + row * 128 + ((ch ^ (row & 7)) << 4)) = v[q];
            }
            fence_proxy_async();                               // generic-proxy stores -> visible to the MMA's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive_count(&full[s], 32);    // the warp's 32 arrivals in one instruction (same count as the cp.async form)
            s += ngrp; if (s >= STAGES) { s -= STAGES; ph ^= 1; }
            cb_c += ngrp; while (cb_c >= p.ga.kb_per_tap) { cb_c -= p.ga.kb_per_tap; ++tp_c; }
            continue;
          }
          if (p.a_mode == OP_K_GATHER) {
            // thread t stages row (m0 + t): 8 chunks of 16 B
            // thread t stages row (m0 + t): 8 chunks of 16 B.  (Measured alternatives, all slower or equal: four whole rows per
            // warp instruction, zero stores for absent rows, TMA tile::gather4 -- the loop is bound by barrier traffic, not by
            // the copies: with the copies AND the B loads knocked out it ran at the same ns per k-block.)
            int tp = tp_c, c0 = cb_c * BK;
            int src_row;
            if (p.ga.idx_in_smem) {
              src_row = sIdx[tp * EPI_THREADS + t];
            } else {
              if (p.ga.mirror_taps) tp = p.ga.ntaps - 1 - tp;
              int m = m0 + t;
              src_row = (m < p.M) ? p.ga.idx[(int64_t)tp * p.ga.idx_stride + m] : -1;
            }
            // (eight lanes per row with cp.async -- four rows per warp instruction -- measured SLOWER than this form: LDGSTS is
            // handled per thread, 16 bytes per request, whatever the addresses)
            const __nv_bfloat16* src = p.ga.base + (int64_t)max(src_row, 0) * p.ga.row_elems + c0;
            uint8_t* dst = sA + s * A_BYTES + t * 128;
            if (!(p.debug & 64)) {
#pragma unroll
              for (int c = 0; c < 8; ++c) cp_async_16(dst + ((c ^ (t & 7)) << 4), src + c * 8, src_row >= 0);
            }
          } else if (p.a_mode == OP_MN_GATHER) {
            stage_mn_gather<BM>(sA + s * A_BYTES, p.ga, a_row, m0, p.M, t);
          }
          if (b_gather) stage_mn_gather<BN>(sB + s * B_BYTES, p.gb, b_row, n0, p.N, t);
        }
        if (p.async_arrive) {
          // the barrier tracks this thread's copies: it arrives when they land, the thread moves on to the next free stage
          // (with commit/wait_group the slack left for the MMA round trip was STAGES - LAG = 2 stages: ~650 ns per k-block)
          if (i < num_kb) {
            cp_async_mbar_arrive_noinc(&full[s]);
            s += ngrp; if (s >= STAGES) { s -= STAGES; ph ^= 1; }
            cb_c += ngrp; while (cb_c >= p.ga.kb_per_tap && p.a_mode == OP_K_GATHER) { cb_c -= p.ga.kb_per_tap; ++tp_c; }
          }
          continue;
        }
        if (i < num_kb) {
          s += 1; if (s >= STAGES) { s -= STAGES; ph ^= 1; }       // (ngrp == 1 in this mode)
          cb_c += 1; while (cb_c >= p.ga.kb_per_tap && p.a_mode == OP_K_GATHER) { cb_c -= p.ga.kb_per_tap; ++tp_c; }
        }
        cp_async_commit();
        if (i >= LAG) {
          if (LAG == 4) cp_async_wait<4>(); else if (LAG == 3) cp_async_wait<3>(); else if (LAG == 2) cp_async_wait<2>(); else cp_async_wait<1>();
          fence_proxy_async();
          __syncwarp();                                  // every lane's copies have landed and are fenced: ONE arrival per warp
          if (lane == 0) mbar_arrive(&full[s_lag]);      // (128 per-thread arrivals serialised on the barrier word)
          if (++s_lag == STAGES) s_lag = 0;
        }
      }
    }
  }

  // ===================== epilogue (warps 2-5) =====================
  // TMEM -> registers -> (bias, activation, dropout, residual) -> swizzled shared-memory tile -> TMA store
  // (or TMA reduce-add for the split-K weight gradients).  The tile reuses the pipeline buffers, which are idle once
  // the last MMA has retired; TMA clips rows >= M and columns >= N.
  if (warp >= 2 && num_kb > 0 && !(p.debug & 1)) {
    const int lane_grp = warp & 3;                     // TMEM lanes [32*lane_grp, +32) are visible to this warp
    const int half = (warp - 2) >> 2;                  // warps w and w+4 share a lane group and split the columns
    const int r = lane_grp * 32 + lane;                // row inside the tile
    const int row = m0 + r;
    if (!(p.debug & 2)) {
      // one poller per warp (256 threads spinning on try_wait for the whole mainloop slowed every barrier hand-off of the pipeline)
      if (lane == 0) mbar_wait(tmem_full, 0);
      __syncwarp();
      tc_fence_after();
    }
    if (threadIdx.x == 64) TS(4);
    if (p.rowsum_a && half == 0 && blockIdx.y == 0) {        // one n-tile column of CTAs reports the sums (every split adds its part)
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)BN, v);
      tmem_ld_wait();
      const float sum = __uint_as_float(v[0]);
      if (row < p.M && sum != 0.f) atomicAdd(&p.rowsum_a[row], sum);
    }
    if (p.residual && !(p.debug & 2)) { if (lane == 0) mbar_wait(res_full, 0); __syncwarp(); }
    const bool add_res = p.residual != nullptr && p.act != ACT_GELU_GRAD;
    uint8_t* sOut2 = smem + (size_t)BM * BN * 2;       // pre-activation tile (bf16 outputs only), behind the output tile
    const uint32_t thr = drop_thr(p.drop_p);
    const bool has_drop = p.drop_p > 0.f;
    const float keep_scale = has_drop ? drop_scale(thr) : 1.f;
    const float half_ks = 0.5f * keep_scale;
    const uint32_t thr4 = thr * 0x01010101u;
    const bool mask_in_float = has_drop && (add_res || p.out_fp32);   // otherwise the mask is ANDed into the packed bf16 pairs
    uint8_t* sOut = smem;                              // [BN/box_cols][128 rows][128 B], SWIZZLE_128B
    const uint32_t sw = (uint32_t)(r & 7);
#pragma unroll 1
    for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
      uint32_t v[32];
      if (!(p.debug & 8)) {
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (uint32_t)(c0 + j + r);
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
      if (p.out2) {                                    // pre-activation copy (bf16): same swizzled layout as the output tile
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
        const uint64_t o0 = (uint64_t)((int64_t)row * p.ldo + nbase);   // multiple of 4: four columns share one hash
#pragma unroll
        for (int j = 0; j < 8; ++j) dm[j] = __vcmpgeu4(drop_hash4(p.seed, p.drop_stream, o0 + 4 * j), thr4);   // 0xFF keep / 0x00 drop
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) dm[j] = 0xFFFFFFFFu;
      }
      // activation (the keep-scale of the dropout is folded into it)
      if (p.act == ACT_GELU_GRAD) {                    // f *= gelu'(aux), aux = bf16 tile loaded through the residual map
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
      if (mask_in_float) {                             // a residual (or an fp32 store) follows: zero the dropped values now
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int q = 0; q < 4; ++q) f[4 * j + q] = ((dm[j] >> (8 * q)) & 1u) ? f[4 * j + q] : 0.f;
      }
      if (p.debug & 16) { if (f[3] == 12345.678f) sBias[0] = f[5]; continue; }
      if (p.out_fp32) {
        // 32 fp32 columns = one 128-byte box row: box index c0/32, 8 chunks of 16 B
        uint8_t* box = sOut + (c0 >> 5) * 16384 + r * 128;
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
        // 32 bf16 columns = half of a 128-byte box row: box index c0/64, chunks (c0%64)/8 .. +3
        uint8_t* box = sOut + (c0 >> 6) * 16384 + r * 128;
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
          if (has_drop && !mask_in_float) {            // bytes (keep ? 0xFF : 0) of four columns -> 16-bit lanes of the bf16 pairs
            w.x &= __byte_perm(dm[2 * ch], 0, 0x1100); w.y &= __byte_perm(dm[2 * ch], 0, 0x3322);
            w.z &= __byte_perm(dm[2 * ch + 1], 0, 0x1100); w.w &= __byte_perm(dm[2 * ch + 1], 0, 0x3322);
          }
          *(uint4*)(box + (((cb + ch) ^ sw) << 4)) = w;
        }
      }
    }
    fence_proxy_async();                               // generic-proxy smem writes -> visible to the TMA (async proxy)
    asm volatile("bar.sync 1, 256;" ::: "memory");     // the eight epilogue warps
    if (threadIdx.x == 64) TS(5);
    if (threadIdx.x == 64 && !(p.debug & 32)) {
      const int box_cols = p.out_fp32 ? 32 : 64;
      const int ncol0 = n0 + (int)((int64_t)tap * p.tap_out_stride);      // per-tap column offset (conv weight gradient)
      for (int j = 0; j < BN / box_cols; ++j) {
        if (n0 + j * box_cols >= p.N) break;
        if (p.atomic) tma_reduce_add_2d(&map_out, sOut + j * 16384, ncol0 + j * box_cols, m0);
        else tma_store_2d(&map_out, sOut + j * 16384, ncol0 + j * box_cols, m0);
        if (p.out2) tma_store_2d(&map_out2, sOut2 + j * 16384, ncol0 + j * box_cols, m0);
      }
      if (p.debug & 4) asm volatile("cp.async.bulk.commit_group;" ::: "memory"); else tma_store_commit_wait();
      TS(6);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
  if (threadIdx.x == 0) TS(7);
}

// ------------------------------------------------------------------------------------------------ host side
static unsigned long long* g_dbg_ts = nullptr;

struct Operand {
  int mode;
  const __nv_bfloat16* ptr;   // TMA modes: matrix base
  int64_t ld;                 // TMA modes: row stride in elements
  GatherSpec g;               // gather modes
};

static bool force_deep_gather() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LOTUS_TC_DEEP_GATHER"); v = e ? atoi(e) : 0; }   // measured on B200: 5-6 stages cost ~1 ms/step (one 200 KB CTA per SM, nothing co-resident), so off
  return v != 0;
}

static int launch(const Operand& A, const Operand& B, Params p, int BN, int splits, cudaStream_t st, int taps = 1,
                  int max_kb_hint = -1) {
  if (p.M <= 0 || p.N <= 0) return 0;
  CUtensorMap ma, mb, mo, mr, mo2;
  memset(&mo2, 0, sizeof(mo2));
  memset(&ma, 0, sizeof(ma));
  memset(&mb, 0, sizeof(mb));
  memset(&mo, 0, sizeof(mo));
  memset(&mr, 0, sizeof(mr));
  p.a_mode = A.mode; p.b_mode = B.mode; p.ga = A.g; p.gb = B.g;
  int rc = 0;
  if (A.mode == OP_K_TMA) rc = make_map(&ma, A.ptr, p.M, p.K, A.ld, BK, BM);
  else if (A.mode == OP_K_TMAGATHER) rc = make_map(&ma, A.g.base, A.g.src_rows, A.g.row_elems, A.g.row_elems, BK, 1);   // box {64 columns, 1 row}
  else if (A.mode == OP_MN_TMA) rc = make_map(&ma, A.ptr, p.K, p.M, A.ld, 64, BK);
  if (rc) return rc;
  if (B.mode == OP_K_TMA) rc = make_map(&mb, B.ptr, p.N, p.K, B.ld, BK, BN);
  else if (B.mode == OP_MN_TMA) rc = make_map(&mb, B.ptr, p.K, p.N, B.ld, 64, BK);
  if (rc) return rc;
  {
    // output (and residual) tile maps: logical [M rows][ld cols]; conv weight gradients address tap d at column offset d*tap_out_stride
    int box_cols = p.out_fp32 ? 32 : 64;
    int64_t out_cols = p.tap_out_stride ? p.ldo : p.N;
    rc = make_map(&mo, p.out, p.M, out_cols, p.ldo, box_cols, BM, p.out_fp32 != 0);
    if (rc) return rc;
    if (p.residual) rc = make_map(&mr, p.residual, p.M, p.N, p.ldo, box_cols, BM, p.out_fp32 != 0);
    if (rc) return rc;
    if (p.out2) rc = make_map(&mo2, p.out2, p.M, out_cols, p.ldo, box_cols, BM, false);
    if (rc) return rc;
  }
  int num_kb = (p.K + BK - 1) / BK;
  splits = max(1, min(splits, num_kb));
  int per = (num_kb + splits - 1) / splits;
  if (!p.k_dev) splits = (num_kb + per - 1) / per;
  p.kb_begin_stride = per;
  p.splits = splits;
  int kb_per_cta = max_kb_hint > 0 ? max_kb_hint : per;
  p.stages = kb_per_cta >= 4 ? 4 : (kb_per_cta >= 3 ? 3 : 2);
  {
    static int force_stages = -1;                       // dev knob: LOTUS_TC_STAGES=2..4
    if (force_stages < 0) { const char* e = getenv("LOTUS_TC_STAGES"); force_stages = e ? atoi(e) : 0; }
    if (force_stages >= 2 && force_stages <= 4) p.stages = min(force_stages, max(2, kb_per_cta));
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("LOTUS_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
    p.dbg_ts = g_dbg_ts;
  }
  // dev knob (LOTUS_TC_DEEP_GATHER=1): 5-6 stage pipeline for the gathered conv operand (3-4 k-blocks of cp.async in flight)
  if (A.mode == OP_K_GATHER && BN == 128 && kb_per_cta >= 6 && force_deep_gather()) p.stages = p.residual ? 5 : 6;
  {
    // long reductions: the per-CTA rate is STAGES per (load latency + hand-offs) ~ 1.5 us, i.e. ~380 ns per k-block at 4 stages
    // whatever the tile does; with >= 12 k-blocks per CTA take 6 stages (5 with a residual tile) -- LOTUS_TC_DEEP=0 disables
    static int deep = -1;
    if (deep < 0) { const char* e = getenv("LOTUS_TC_DEEP"); deep = e ? atoi(e) : 0; }
    if (deep && kb_per_cta >= 12 && !p.k_dev) p.stages = (BN == 128) ? (p.residual ? 5 : 6) : 6;
  }
  {
    static int aa = -1;                                 // dev knob: LOTUS_TC_ASYNC_ARRIVE=0 restores commit/wait_group signalling
    if (aa < 0) { const char* e = getenv("LOTUS_TC_ASYNC_ARRIVE"); aa = e ? atoi(e) : 1; }
    p.async_arrive = aa;
    static int lg = -1;                                 // dev knob: LOTUS_TC_LDG_GATHER=1 gathers through registers (8 lanes per row)
    if (lg < 0) { const char* e = getenv("LOTUS_TC_LDG_GATHER"); lg = e ? atoi(e) : 2; }   // 0 cp.async, 1 LDG + STS, 2 by row count (see the kernel)
    p.ldg_gather = lg == 2 ? (p.M <= 65536 ? 1 : 0) : lg;
  }
  dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN, splits * taps);
  size_t tile_bytes = (size_t)BM * BN * (p.out_fp32 ? 4 : 2);
  int want_pair;
  { static int pc = -1; if (pc < 0) { const char* e = getenv("LOTUS_TC_PAIR"); pc = e ? atoi(e) : 0; } want_pair = pc; }
  size_t pipe = (size_t)p.stages * (BM * BK * 2 + BN * BK * 2);
  while (pipe < tile_bytes * (p.out2 ? 2 : 1)) { ++p.stages; pipe = (size_t)p.stages * (BM * BK * 2 + BN * BK * 2); }   // the output tile reuses the pipeline
  size_t smem = pipe + 256 + 2048 + (size_t)BN * 4 + (p.residual ? tile_bytes : 0) + (p.rowsum_a ? 2048 : 0);
  if ((A.mode == OP_K_GATHER || A.mode == OP_K_TMAGATHER) && A.g.ntaps <= 32) {
    // neighbour-table tile in shared memory; keep two 64-wide CTAs per SM (<= ~113 KB each) by giving up one pipeline stage
    size_t idx_bytes = (size_t)A.g.ntaps * EPI_THREADS * 4;
    if (BN == 64 && smem + idx_bytes > 113 * 1024 && p.stages > 2 && (size_t)(p.stages - 1) * (BM * BK * 2 + BN * BK * 2) >= tile_bytes) {
      --p.stages;
      pipe = (size_t)p.stages * (BM * BK * 2 + BN * BK * 2);
      smem = pipe + 256 + 2048 + (size_t)BN * 4 + (p.residual ? tile_bytes : 0);
    }
    smem += idx_bytes;
    p.ga.idx_in_smem = 1;
  }
  p.pair_commit = (want_pair && (p.stages % 2 == 0)) ? 1 : 0;
  cudaError_t e;
  if (BN == 128) {
    static bool set128 = false;
    if (!set128) { cudaFuncSetAttribute(k_gemm_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); set128 = true; }
    launch_pdl(k_gemm_tc<128>, grid, dim3(NUM_THREADS), smem, st, ma, mb, mo, mr, mo2, p);
  } else {
    static bool set64 = false;
    if (!set64) { cudaFuncSetAttribute(k_gemm_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); set64 = true; }
    launch_pdl(k_gemm_tc<64>, grid, dim3(NUM_THREADS), smem, st, ma, mb, mo, mr, mo2, p);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) { lotus_set_error("gemm_tc launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

static int pick_bn(int N) { return N <= 64 ? 64 : 128; }
// sparse-conv A operand: cp.async gather by default; LOTUS_TC_TMA_GATHER=1 gathers with the TMA unit instead (tile::gather4:
// correct, and frees the LSU, but the TMA unit retires one 128-byte row per ~10 ns, half the cp.async rate)
static int conv_gather_mode(const void* src, int K3) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LOTUS_TC_TMA_GATHER"); v = e ? atoi(e) : 0; }   // measured: gather4 sustains ~1290 ns per 128-row k-block, cp.async ~610
  return (v && K3 <= 32 && ((uintptr_t)src & 15) == 0) ? OP_K_TMAGATHER : OP_K_GATHER;
}

}  // namespace tc

using tc::Operand;
using tc::Params;

static Params base_params(int M, int N, int K, void* out, int64_t ldo, int out_fp32) {
  Params p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.out = out; p.ldo = ldo; p.out_fp32 = out_fp32;
  return p;
}

// dw[N,K] (fp32) += dy[M,N]^T @ x[M,K]   (both operands consumed MN-major straight from their row-major storage)
LOTUS_API int lotus_linear_wgrad_bf16(const void* dy, const void* x, float* dw, int M, int N, int K, void* stream_) {
  LOTUS_CHECK_ARG(N % 8 == 0 && K % 8 == 0, "lotus_linear_wgrad_bf16: N=%d, K=%d must be multiples of 8", N, K);
  Params p = base_params(N, K, M, dw, K, 1);
  p.atomic = 1;
  Operand A{tc::OP_MN_TMA, (const __nv_bfloat16*)dy, N, {}}, B{tc::OP_MN_TMA, (const __nv_bfloat16*)x, K, {}};
  int bn = tc::pick_bn(K);
  int tiles = ((N + 127) / 128) * ((K + bn - 1) / bn);
  static int ctas = -1;                                          // CTAs the reduction is cut over: one per SM (measured best at every shape of the step: two per SM doubles the reduce-add traffic onto the same tile); LOTUS_WGRAD_CTAS overrides
  if (ctas < 0) { const char* e = getenv("LOTUS_WGRAD_CTAS"); ctas = e ? atoi(e) : kNumSM; }
  int splits = max(1, ctas / tiles);
  return tc::launch(A, B, p, bn, splits, (cudaStream_t)stream_);
}

// weight and bias gradient in one GEMM: dw[N,K] (fp32) += dy^T x, db[N] (fp32) += dy^T 1 (either may be null)
extern "C" int lotus_colsum_bf16(const void* x, int64_t rows, int C, float* out_accum, void* stream_);
LOTUS_API int lotus_linear_wgrad_bias_bf16(const void* dy, const void* x, float* dw, float* db, int M, int N, int K, void* stream_) {
  if (dw == nullptr) return db ? lotus_colsum_bf16(dy, M, N, db, stream_) : 0;
  LOTUS_CHECK_ARG(N % 8 == 0 && K % 8 == 0, "lotus_linear_wgrad_bias_bf16: N=%d, K=%d must be multiples of 8", N, K);
  Params p = base_params(N, K, M, dw, K, 1);
  p.atomic = 1;
  p.rowsum_a = db;
  Operand A{tc::OP_MN_TMA, (const __nv_bfloat16*)dy, N, {}}, B{tc::OP_MN_TMA, (const __nv_bfloat16*)x, K, {}};
  int bn = tc::pick_bn(K);
  int tiles = ((N + 127) / 128) * ((K + bn - 1) / bn);
  static int ctas = -1;
  if (ctas < 0) { const char* e = getenv("LOTUS_WGRAD_CTAS"); ctas = e ? atoi(e) : kNumSM; }
  int splits = max(1, ctas / tiles);
  return tc::launch(A, B, p, bn, splits, (cudaStream_t)stream_);
}

// submanifold conv forward: y[n,Cout] = sum_{d,ci} x[nbr[d][n], ci] * w[co, d, ci] (+bias); Cin % 64 == 0
LOTUS_API int lotus_subm_conv_fwd_bf16(const void* x, const void* w, const float* bias, void* y, const int32_t* nbr,
                                       int64_t nbr_stride, int n, int Cin, int Cout, int K3, int out_fp32, void* stream_) {
  LOTUS_CHECK_ARG(Cin % 64 == 0, "lotus_subm_conv_fwd_bf16: Cin=%d must be a multiple of 64", Cin);
  Params p = base_params(n, Cout, K3 * Cin, y, Cout, out_fp32);
  p.bias = bias;
  Operand A{tc::conv_gather_mode(x, K3), nullptr, 0, {(const __nv_bfloat16*)x, nbr, nbr_stride, Cin, Cin / 64, 0, K3, 0, n}};
  Operand B{tc::OP_K_TMA, (const __nv_bfloat16*)w, (int64_t)K3 * Cin, {}};
  return tc::launch(A, B, p, tc::pick_bn(Cout), 1, (cudaStream_t)stream_);
}

// dgrad: dx[q,ci] = [residual[q,ci] +] sum_{d,co} dy[nbr[K3-1-d][q], co] * w[co,d,ci], with wt = w re-laid out as [ci][d][co] (bf16)
LOTUS_API int lotus_subm_conv_dgrad_bf16(const void* dy, const void* wt, const void* residual, void* dx, const int32_t* nbr,
                                         int64_t nbr_stride, int n, int Cin, int Cout, int K3, int out_fp32, void* stream_) {
  LOTUS_CHECK_ARG(Cout % 64 == 0, "lotus_subm_conv_dgrad_bf16: Cout=%d must be a multiple of 64", Cout);
  Params p = base_params(n, Cin, K3 * Cout, dx, Cin, out_fp32);
  p.residual = residual;
  Operand A{tc::conv_gather_mode(dy, K3), nullptr, 0, {(const __nv_bfloat16*)dy, nbr, nbr_stride, Cout, Cout / 64, 1, K3, 0, n}};
  Operand B{tc::OP_K_TMA, (const __nv_bfloat16*)wt, (int64_t)K3 * Cout, {}};
  return tc::launch(A, B, p, tc::pick_bn(Cin), 1, (cudaStream_t)stream_);
}

// ---- split-K forms for the deep stages ----
// With n = 976 rows the grid of the plain form is 48 CTAs, each walking 324 k-blocks in series: 140-160 us for a tensor of 1.5 MB on a
// third of the SMs (the 3210- and 12140-row stages launch 104 / 190 CTAs and, with the coalesced gather, no longer gain from a split).  Here the reduction is cut into `splits` ranges
// (grid z), the partial tiles are added into an fp32 scratch [n, N] with TMA reduce-add, and a small pass adds the bias / residual and
// rounds to bf16.  lotus_subm_conv_splits() says how many ranges a shape gets (1 = use the plain entry point, no scratch needed).
static int conv_splits(int n, int N, int K) {
  const int bn = tc::pick_bn(N);
  const int tiles = ((n + 127) / 128) * ((N + bn - 1) / bn);
  const int num_kb = (K + 63) / 64;
  if (tiles * 2 >= kNumSM) return 1;                             // with the coalesced gather only grids under half a wave gain from the split (measured)
  int s = (kNumSM * 3) / max(1, tiles);                          // three CTAs of this kernel fit an SM
  s = min(s, min(8, num_kb / 12));
  return max(1, s);
}
LOTUS_API int lotus_subm_conv_splits(int n, int N, int K) { return conv_splits(n, N, K); }

__global__ void __launch_bounds__(256) k_splitk_finalize(const float* __restrict__ acc, const float* __restrict__ bias,
                                                         const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out,
                                                         int64_t nvec, int C) {
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = *(const float4*)(acc + i * 8), b = *(const float4*)(acc + i * 8 + 4);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const int c = (int)((i * 8) % C);
    if (bias) {
      const float4 p0 = *(const float4*)(bias + c), p1 = *(const float4*)(bias + c + 4);
      v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w; v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
    }
    if (residual) {
      const uint4 q = *(const uint4*)(residual + i * 8);
      const __nv_bfloat162* qh = (const __nv_bfloat162*)&q;
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(qh[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
    }
    uint4 w;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
    w.x = *(uint32_t*)&h0; w.y = *(uint32_t*)&h1; w.z = *(uint32_t*)&h2; w.w = *(uint32_t*)&h3;
    *(uint4*)(out + i * 8) = w;
  }
}
static int splitk_conv(const void* src, const void* wmat, const float* bias, const void* residual, void* out, const int32_t* nbr,
                       int64_t nbr_stride, int n, int Cred, int N, int K3, int mirror, float* scratch, cudaStream_t st) {
  const int K = K3 * Cred, splits = conv_splits(n, N, K);
  LOTUS_CHECK_ARG(scratch != nullptr && N % 8 == 0, "split-K conv: needs an fp32 scratch of n * N floats and N %% 8 == 0 (N=%d)", N);
  LOTUS_CUDA_CALL(cudaMemsetAsync(scratch, 0, (size_t)n * N * 4, st));
  Params p = base_params(n, N, K, scratch, N, 1);
  p.atomic = 1;
  Operand A{tc::conv_gather_mode(src, K3), nullptr, 0, {(const __nv_bfloat16*)src, nbr, nbr_stride, Cred, Cred / 64, mirror, K3, 0, n}};
  Operand B{tc::OP_K_TMA, (const __nv_bfloat16*)wmat, (int64_t)K, {}};
  int rc = tc::launch(A, B, p, tc::pick_bn(N), splits, st);
  if (rc) return rc;
  const int64_t nvec = (int64_t)n * N / 8;
  launch_pdl(k_splitk_finalize, dim3((unsigned)min((int64_t)kNumSM * 8, (nvec + 255) / 256)), dim3(256), 0, st, (const float*)scratch, bias,
             (const __nv_bfloat16*)residual, (__nv_bfloat16*)out, nvec, N);
  LOTUS_CUDA_OK();
  return 0;
}
// forward with the reduction split (bf16 output); scratch: n * Cout floats
LOTUS_API int lotus_subm_conv_fwd_splitk_bf16(const void* x, const void* w, const float* bias, void* y, const int32_t* nbr,
                                              int64_t nbr_stride, int n, int Cin, int Cout, int K3, float* scratch, void* stream_) {
  LOTUS_CHECK_ARG(Cin % 64 == 0, "lotus_subm_conv_fwd_splitk_bf16: Cin=%d must be a multiple of 64", Cin);
  if (n <= 0) return 0;
  return splitk_conv(x, w, bias, nullptr, y, nbr, nbr_stride, n, Cin, Cout, K3, 0, scratch, (cudaStream_t)stream_);
}
// data gradient with the reduction split (bf16 output, optional bf16 residual); scratch: n * Cin floats
LOTUS_API int lotus_subm_conv_dgrad_splitk_bf16(const void* dy, const void* wt, const void* residual, void* dx, const int32_t* nbr,
                                                int64_t nbr_stride, int n, int Cin, int Cout, int K3, float* scratch, void* stream_) {
  LOTUS_CHECK_ARG(Cout % 64 == 0, "lotus_subm_conv_dgrad_splitk_bf16: Cout=%d must be a multiple of 64", Cout);
  if (n <= 0) return 0;
  return splitk_conv(dy, wt, nullptr, residual, dx, nbr, nbr_stride, n, Cout, Cin, K3, 1, scratch, (cudaStream_t)stream_);
}

// wgrad: dw[co, d, ci] (fp32) += sum_p dy[p,co] * x[nbr[d][p], ci]
LOTUS_API int lotus_subm_conv_wgrad_bf16(const void* dy, const void* x, float* dw, const int32_t* nbr, int64_t nbr_stride,
                                         int n, int Cin, int Cout, int K3, void* stream_) {
  LOTUS_CHECK_ARG(Cin % 64 == 0 && Cout % 8 == 0, "lotus_subm_conv_wgrad_bf16: Cin=%d (x64), Cout=%d (x8)", Cin, Cout);
  int bn = tc::pick_bn(Cin);
  int tiles = ((Cout + 127) / 128) * ((Cin + bn - 1) / bn);
  int splits = max(1, (kNumSM * 2) / (tiles * K3) + 1);
  for (int d = 0; d < K3; ++d) {
    Params p = base_params(Cout, Cin, n, dw + (int64_t)d * Cin, (int64_t)K3 * Cin, 1);
    p.atomic = 1;
    Operand A{tc::OP_MN_TMA, (const __nv_bfloat16*)dy, Cout, {}};
    Operand B{tc::OP_MN_GATHER, nullptr, 0, {(const __nv_bfloat16*)x, nbr + (int64_t)d * nbr_stride, 0, Cin, 0, 0, K3}};
    int rc = tc::launch(A, B, p, bn, splits, (cudaStream_t)stream_);
    if (rc) return rc;
  }
  return 0;
}

// wgrad over the compact pair lists (lotus_pair_lists): dw[co, d, ci] (fp32) += sum_j dy[pairs_out[d][j], co] * x[pairs_in[d][j], ci]
// One launch for all taps: grid.z = K3 * splits; the per-tap reduction extent is read on the device (pair_cnt).
LOTUS_API int lotus_subm_conv_wgrad_pairs_bf16(const void* dy, const void* x, float* dw, const int32_t* pairs_out,
                                               const int32_t* pairs_in, const int32_t* pair_cnt, int64_t pair_stride,
                                               int n_upper, int Cin, int Cout, int K3, void* stream_) {
  LOTUS_CHECK_ARG(Cin % 8 == 0 && Cout % 8 == 0, "lotus_subm_conv_wgrad_pairs_bf16: Cin=%d, Cout=%d must be multiples of 8", Cin, Cout);
  int bn = tc::pick_bn(Cin);
  int tiles = ((Cout + 127) / 128) * ((Cin + bn - 1) / bn);
  int max_kb = (n_upper + tc::BK - 1) / tc::BK;
  // CTAs the (tap, tile) reductions are cut over.  Unlike the dense weight gradient this kernel is bound by the request rate of its
  // gathers, so the sparse 96k-row stage wants MORE CTAs in flight (measured, tools/conv_bench.py: 5 per SM 56 / 114 us, 3 per SM
  // 79 / 135 us at C = 64 / 128), the deep stages fewer (less reduce-add traffic).  LOTUS_CONV_WGRAD_CTAS overrides.
  static int ctas_env = -1;
  if (ctas_env < 0) { const char* e = getenv("LOTUS_CONV_WGRAD_CTAS"); ctas_env = e ? atoi(e) : 0; }
  const int ctas = ctas_env > 0 ? ctas_env : (n_upper > 65536 ? kNumSM * 5 : (n_upper > 8192 ? kNumSM * 3 : kNumSM * 2));
  int splits = max(1, min(max_kb, ctas / max(1, tiles * K3) + 1));
  Params p = base_params(Cout, Cin, n_upper, dw, (int64_t)K3 * Cin, 1);
  p.atomic = 1;
  p.k_dev = pair_cnt;
  p.tap_idx_stride = pair_stride;
  p.tap_out_stride = Cin;
  Operand A{tc::OP_MN_GATHER, nullptr, 0, {(const __nv_bfloat16*)dy, pairs_out, 0, Cout, 0, 0, K3}};
  Operand B{tc::OP_MN_GATHER, nullptr, 0, {(const __nv_bfloat16*)x, pairs_in, 0, Cin, 0, 0, K3}};
  return tc::launch(A, B, p, bn, splits, (cudaStream_t)stream_, K3, 4);
}

// dev tool: device buffer (8 x uint64 per CTA) receiving %globaltimer stamps of the next GEMM launches; null disables
LOTUS_API int lotus_tc_debug_buffer(void* buf) { tc::g_dbg_ts = (unsigned long long*)buf; return 0; }
