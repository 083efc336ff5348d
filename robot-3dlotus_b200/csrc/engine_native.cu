// Native executor of one policy step: the whole forward (+ loss + backward) of SimplePolicyPTV3CA issued from
// C++ in one call, so the ~900 kernel launches of a training step cost microseconds of host time instead of
// a Python/autograd round trip each.  It replays exactly the op sequence of the Python engine
// (robot-3dlotus_b200/engine.py, itself following reference models/simple_policy_ptv3.py:225-306,
// models/PointTransformerV3/model_ca.py:383-412, model.py:659-680,713-790,817-828) through the same launchers.
//
// Memory: a caller-provided arena (bump allocator); activations saved for the backward live until the step
// ends, backward temporaries are released per block.  Gradients: parameter gradients ACCUMULATE into the
// caller's fp32 sinks (the trainer's flat buffer); activation gradients flow through one "D" buffer per stage
// that is updated in place along the residual stream (no gradient-accumulation add kernels).
//
// Scope: the production topology (CA backbone, depth 1 per stage, plain BN/LN, heatmap_disc/euler_disc head,
// text-only context).  Other configurations run through the Python engine.
#include "common.cuh"
#include "../../include/lotus.h"
#include <vector>
#include <type_traits>
#include <stdlib.h>

namespace eng {

struct Arena {
  char* base; size_t off, cap, peak; bool dry, overflow;
  void* alloc(size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    void* p = base + off;
    off += bytes;
    if (off > peak) peak = off;
    if (!dry && off > cap) overflow = true;      // nothing is launched once the arena is exhausted
    return p;
  }
};

struct Lin { int w, b, out, in; };
struct Norm { int w, b, C; };
struct BNorm { int w, b, C, buf; };   // buf: index into the running-stat pointer array (mean = 2*buf, var = 2*buf+1)
struct BlockP { int conv_w, conv_b; Lin cpe1; Norm cpe2, n1; Lin qkv, proj; Norm qn, kn, n2; Lin fc1, fc2; int C, H; };
struct CaP { Norm n1; Lin q, kv, proj; Norm qn, kn, n2; Lin fc1, fc2; int C, H; };
struct PoolP { Lin proj; BNorm bn; };
struct UnpoolP { Lin proj; BNorm bn; Lin skip; BNorm bn_skip; };

struct StageInfo {
  int n, n_pad, n_patches, max_cloud; int64_t stride;
  const int32_t *nbr3, *pairs_out, *pairs_in, *pair_cnt, *gather_idx, *slot_of_point, *cu_seqlens, *cloud_off, *cluster, *seg_ptr,
      *member, *nbr5;
};

struct Ctx {
  cudaStream_t st; void* stv; bool bf, train, dry; int es; Arena ar; int rc;
  cudaStream_t side; void* sidev; bool side_busy;            // weight-gradient stream (null: everything on `st`)
  uint64_t seed; uint32_t drop_next; float attn_drop, proj_drop, head_drop;
  void* const* P; void* const* G; void* const* BN;          // fp32 params, grad sinks (may be null), BN running stats
  std::vector<void*> WB, WT;                                 // bf16 staged weights (by param index)
  int B, S, K, ctxC, qk_norm, Lt, max_tok; const int32_t* tok_off;
};

static long long g_launcher_calls = 0;   // launcher invocations of the last non-dry step (each launches >= 1 kernel)
// every kernel-launching C-ABI call of the step goes through CALL: in profile mode (lotus_policy_step_profile) each call is
// bracketed by CUDA events on the launching stream and its scalar / pointer arguments are recorded, so bench.py gets per-call
// device times of the NATIVE path (no Python between launches) together with the shapes the byte / FLOP model needs
struct ProfRec { const char* name; cudaEvent_t e0, e1; int nargs; int64_t a[48]; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_events;
static size_t g_prof_ev_next = 0;
static cudaEvent_t prof_event() {
  if (g_prof_ev_next == g_prof_events.size()) { cudaEvent_t e; cudaEventCreate(&e); g_prof_events.push_back(e); }
  return g_prof_events[g_prof_ev_next++];
}
template <typename T> static inline int64_t prof_arg(T v) {
  if constexpr (std::is_null_pointer_v<T>) return 0;
  else if constexpr (std::is_pointer_v<T>) return (int64_t)(uintptr_t)v;
  else if constexpr (std::is_floating_point_v<T>) return (int64_t)(v != 0);
  else return (int64_t)v;
}
template <typename F, typename... A> static int prof_call(cudaStream_t st, const char* name, F fn, A... args) {
  if (!g_prof_on) return fn(args...);
  ProfRec r; r.name = name; r.nargs = (int)sizeof...(A) < 48 ? (int)sizeof...(A) : 48;
  int i = 0; ((r.a[i < 48 ? i : 47] = prof_arg(args), ++i), ...);
  r.e0 = prof_event(); r.e1 = prof_event();
  cudaEventRecord(r.e0, st);
  int rc = fn(args...);
  cudaEventRecord(r.e1, st);
  g_prof.push_back(r);
  return rc;
}
#define CALL(fn, ...) do { if (!c.dry && c.rc == 0 && !c.ar.overflow) { ++g_launcher_calls; int _r = prof_call(c.st, #fn, fn, __VA_ARGS__); if (_r) c.rc = _r; } } while (0)

// Weight / bias gradients do not sit on the backward's critical path (the activation-gradient chain): they run on a
// second stream so the small deep-stage kernels overlap.  fork: the side stream waits for everything issued so far on the
// main stream; join: the main stream waits for the side stream (before arena memory the side kernels read is reused, or
// a buffer they read is modified in place, and at the end of the step).
static cudaStream_t g_side = nullptr;
static cudaEvent_t g_ev_fork = nullptr, g_ev_join = nullptr;
static inline bool live(Ctx& c) { return !c.dry && c.rc == 0 && !c.ar.overflow; }
static void fork_side(Ctx& c) {
  if (!c.side || !live(c)) return;
  cudaEventRecord(g_ev_fork, c.st);
  cudaStreamWaitEvent(c.side, g_ev_fork, 0);
  c.side_busy = true;
}
static void join_side(Ctx& c) {
  if (!c.side || !c.side_busy || c.dry) return;
  cudaEventRecord(g_ev_join, c.side);
  cudaStreamWaitEvent(c.st, g_ev_join, 0);
  c.side_busy = false;
}
static inline void* gstream(Ctx& c) { return c.side ? c.sidev : c.stv; }   // stream of the weight-gradient kernels

static inline float* Pf(Ctx& c, int i) { return i < 0 ? nullptr : (float*)c.P[i]; }
static inline float* Gf(Ctx& c, int i) { return (i < 0 || !c.G) ? nullptr : (float*)c.G[i]; }
static inline void* act(Ctx& c, int64_t rows, int64_t cols) { return c.ar.alloc((size_t)rows * cols * c.es); }
static inline float* f32(Ctx& c, int64_t n) { return (float*)c.ar.alloc((size_t)n * 4); }
static inline void zero(Ctx& c, void* p, size_t bytes) { if (!c.dry && c.rc == 0 && !c.ar.overflow) cudaMemsetAsync(p, 0, bytes, c.st); }
static inline uint32_t drop_id(Ctx& c, float p) { return (c.train && p > 0.f) ? ++c.drop_next : 0; }
static inline float drop_p(Ctx& c, float p, uint32_t id) { return id ? p : 0.f; }

// ------------------------------------------------------------------------------------------------ primitive wrappers
static void linear_fwd(Ctx& c, const void* x, int M, const Lin& l, const void* residual, void* y, float dp, uint32_t did,
                       bool out_fp32 = false, bool x_bf = true) {
  if (c.bf && x_bf)
    CALL(lotus_linear_fwd_bf16, x, c.WB[l.w], Pf(c, l.b), residual, y, nullptr, M, l.out, l.in, 0, out_fp32 ? 1 : 0, drop_p(c, dp, did),
                               c.seed, did, c.stv);
  else
    CALL(lotus_linear_fwd_f32, (const float*)x, Pf(c, l.w), Pf(c, l.b), (const float*)residual, (float*)y, M, l.out, l.in, 0, nullptr,
                              drop_p(c, dp, did), c.seed, did, c.stv);
}
// dy: gradient of the layer output AFTER its dropout (if any) has been applied to dy by the caller.
// dx = [dx_residual +] dy @ W ; dW += dy^T x ; db += colsum(dy).   tc: operands are bf16 (tensor-core path)
static void linear_bwd(Ctx& c, const void* dy, const void* x, int M, const Lin& l, void* dx, const void* dx_residual, bool tc) {
  fork_side(c);                                  // dy (and x) are complete on the main stream at this point
  void* gs = gstream(c);
  if (tc) {
    // the bias gradient rides in the weight-gradient GEMM (db = dy^T 1: one extra MMA per K step against a tile of ones), on the side stream
    if (Gf(c, l.w) || Gf(c, l.b)) CALL(lotus_linear_wgrad_bias_bf16, dy, x, Gf(c, l.w), Gf(c, l.b), M, l.out, l.in, gs);
    if (dx) CALL(lotus_linear_dgrad_bias_bf16, dy, c.WT[l.w], dx_residual, dx, nullptr, M, l.out, l.in, 0, c.stv);
  } else {
    if (Gf(c, l.w)) CALL(lotus_linear_wgrad_f32, (const float*)dy, (const float*)x, Gf(c, l.w), M, l.out, l.in, nullptr, gs);
    if (Gf(c, l.b)) CALL(lotus_colsum_f32, (const float*)dy, M, l.out, Gf(c, l.b), gs);
    if (dx) CALL(lotus_linear_dgrad_f32, (const float*)dy, Pf(c, l.w), (const float*)dx_residual, (float*)dx, M, l.out, l.in, c.stv);
  }
}
// apply a dropout mask to a gradient (out of place into the arena) -- identity when the site is inactive
static const void* drop_grad(Ctx& c, const void* dy, int64_t n, float p, uint32_t did, bool bf) {
  if (!did) return dy;
  void* o = c.ar.alloc((size_t)n * (bf ? 2 : 4));
  if (bf) CALL(lotus_dropout_bf16, dy, o, n, p, c.seed, did, c.stv);
  else CALL(lotus_dropout_f32, (const float*)dy, (float*)o, n, p, c.seed, did, c.stv);
  return o;
}
static void ln_fwd(Ctx& c, const void* x, void* y, const void* residual, const Norm& nm, float* mean, float* rstd, int64_t rows, float eps) {
  if (c.bf) CALL(lotus_layernorm_fwd_bf16, x, 1, nm.C, 0, y, 1, nm.C, 0, residual, Pf(c, nm.w), Pf(c, nm.b), mean, rstd, rows, nm.C, eps, c.stv);
  else CALL(lotus_layernorm_fwd_f32, (const float*)x, 1, nm.C, 0, (float*)y, 1, nm.C, 0, (const float*)residual, Pf(c, nm.w), Pf(c, nm.b), mean, rstd, rows, nm.C, eps, c.stv);
}
// masked / mask_did (optional, bf16 mode): also write dropout_mask(dx) for the NEXT layer's output dropout site (saves the separate
// masking pass over the residual-stream gradient; bit-identical to lotus_dropout_bf16 applied to the stored dx)
static void ln_bwd(Ctx& c, const void* dy, const void* x, void* dx, const Norm& nm, const float* mean, const float* rstd, int64_t rows, int accumulate,
                   void* masked = nullptr, uint32_t mask_did = 0) {
  if (c.bf && masked && mask_did) {
    CALL(lotus_layernorm_bwd_mask_bf16, dy, x, dx, masked, Pf(c, nm.w), mean, rstd, Gf(c, nm.w), Gf(c, nm.b), rows, nm.C, accumulate,
                                       c.proj_drop, c.seed, mask_did, c.stv);
    return;
  }
  if (c.bf) CALL(lotus_layernorm_bwd_bf16, dy, 0, 1, nm.C, 0, x, 1, nm.C, 0, dx, 1, nm.C, 0, Pf(c, nm.w), mean, rstd, Gf(c, nm.w), Gf(c, nm.b), rows, nm.C, accumulate, c.stv);
  else CALL(lotus_layernorm_bwd_f32, (const float*)dy, 1, nm.C, 0, (const float*)x, 1, nm.C, 0, (float*)dx, 1, nm.C, 0, Pf(c, nm.w), mean, rstd, Gf(c, nm.w), Gf(c, nm.b), rows, nm.C, accumulate, c.stv);
}
// per-head LayerNorm over Dh of a strided view (rows = N*H): x row (n,h) at n*x_outer + h*Dh -> y contiguous (N*H, Dh)
static void head_ln_fwd(Ctx& c, const void* x, int64_t x_outer, void* y, const Norm& nm, float* mean, float* rstd, int64_t N, int H) {
  int D = nm.C;
  if (c.bf) CALL(lotus_layernorm_fwd_bf16, x, H, x_outer, D, y, 1, D, 0, nullptr, Pf(c, nm.w), Pf(c, nm.b), mean, rstd, N * H, D, 1e-6f, c.stv);
  else CALL(lotus_layernorm_fwd_f32, (const float*)x, H, x_outer, D, (float*)y, 1, D, 0, nullptr, Pf(c, nm.w), Pf(c, nm.b), mean, rstd, N * H, D, 1e-6f, c.stv);
}
// dy fp32 contiguous (N*H, Dh) -> dx strided like x
static void head_ln_bwd(Ctx& c, const float* dy, const void* x, int64_t x_outer, void* dx, const Norm& nm, const float* mean,
                        const float* rstd, int64_t N, int H) {
  int D = nm.C;
  if (c.bf) CALL(lotus_layernorm_bwd_bf16, dy, 1, 1, D, 0, x, H, x_outer, D, dx, H, x_outer, D, Pf(c, nm.w), mean, rstd, Gf(c, nm.w), Gf(c, nm.b), N * H, D, 0, c.stv);
  else CALL(lotus_layernorm_bwd_f32, dy, 1, D, 0, (const float*)x, H, x_outer, D, (float*)dx, H, x_outer, D, Pf(c, nm.w), mean, rstd, Gf(c, nm.w), Gf(c, nm.b), N * H, D, 0, c.stv);
}
static void pointwise(Ctx& c, const void* x, const void* a, void* y, int64_t n, int op, float dp, uint32_t did, bool bf) {
  if (bf) CALL(lotus_pointwise_bf16, x, a, y, n, op, 0.02f, drop_p(c, dp, did), c.seed, did, c.stv);
  else CALL(lotus_pointwise_f32, (const float*)x, (const float*)a, (float*)y, n, op, 0.02f, drop_p(c, dp, did), c.seed, did, c.stv);
}

struct BnSave { const void* x; float *mean, *rstd; int64_t rows; bool x_fp32; };
static void bn_fwd(Ctx& c, const void* x, bool x_fp32, void* y, int64_t rows, const BNorm& b, BnSave& sv) {
  sv.x = x; sv.rows = rows; sv.x_fp32 = x_fp32;
  sv.mean = f32(c, b.C); sv.rstd = f32(c, b.C);
  float* scratch = f32(c, 2 * b.C);
  float* rm = (float*)c.BN[2 * b.buf]; float* rv = (float*)c.BN[2 * b.buf + 1];
  if (c.bf) CALL(lotus_batchnorm_fwd_bf16, x, x_fp32 ? 1 : 0, y, rows, b.C, Pf(c, b.w), Pf(c, b.b), rm, rv, sv.mean, sv.rstd, scratch, c.train ? 1 : 0, 1e-3f, 0.01f, 1, c.stv);
  else CALL(lotus_batchnorm_fwd_f32, (const float*)x, (float*)y, rows, b.C, Pf(c, b.w), Pf(c, b.b), rm, rv, sv.mean, sv.rstd, scratch, c.train ? 1 : 0, 1e-3f, 0.01f, 1, c.stv);
}
static void bn_bwd(Ctx& c, const void* dy, void* dx, const BNorm& b, const BnSave& sv) {
  float* dg = f32(c, 2 * b.C); float* db = dg + b.C;
  if (c.bf && c.train && b.C % 8 == 0 && b.C <= 1024) {          // one launch: sums, dx, and the adds into the gradient accumulators
    CALL(lotus_batchnorm_bwd_acc_bf16, dy, sv.x, sv.x_fp32 ? 1 : 0, sv.rows, b.C, Pf(c, b.w), Pf(c, b.b), sv.mean, sv.rstd, dx, dg, db, Gf(c, b.w), Gf(c, b.b), 1, c.stv);
    return;
  }
  zero(c, dg, (size_t)2 * b.C * 4);
  if (c.bf) CALL(lotus_batchnorm_bwd_bf16, dy, sv.x, sv.x_fp32 ? 1 : 0, sv.rows, b.C, Pf(c, b.w), Pf(c, b.b), sv.mean, sv.rstd, dx, dg, db, c.train ? 1 : 0, 1, c.stv);
  else CALL(lotus_batchnorm_bwd_f32, (const float*)dy, (const float*)sv.x, sv.rows, b.C, Pf(c, b.w), Pf(c, b.b), sv.mean, sv.rstd, (float*)dx, dg, db, c.train ? 1 : 0, 1, c.stv);
  if (Gf(c, b.w)) CALL(lotus_pointwise_f32, dg, Gf(c, b.w), Gf(c, b.w), b.C, 4, 0.f, 0.f, 0, 0, c.stv);
  if (Gf(c, b.b)) CALL(lotus_pointwise_f32, db, Gf(c, b.b), Gf(c, b.b), b.C, 4, 0.f, 0.f, 0, 0, c.stv);
}

// ------------------------------------------------------------------------------------------------ MLP tail shared by Block / CABlock
struct MlpSave { void *x2, *h2, *u, *g; float *mean, *rstd; uint32_t d_act, d_out; };
static void* mlp_fwd(Ctx& c, void* x2, int n, const Norm& n2, const Lin& fc1, const Lin& fc2, MlpSave& sv) {
  int C = n2.C;
  sv.x2 = x2; sv.mean = f32(c, n); sv.rstd = f32(c, n);
  sv.h2 = act(c, n, C);
  ln_fwd(c, x2, sv.h2, nullptr, n2, sv.mean, sv.rstd, n, 1e-5f);
  sv.u = act(c, n, fc1.out);
  sv.g = act(c, n, fc1.out);
  sv.d_act = drop_id(c, c.proj_drop);
  if (c.bf) {
    // GELU + dropout fused into the GEMM epilogue; the pre-activation is a second TMA store (only kept when a backward follows)
    CALL(lotus_linear_fwd_bf16, sv.h2, c.WB[fc1.w], Pf(c, fc1.b), nullptr, sv.g, c.G ? sv.u : nullptr, n, fc1.out, fc1.in, 1, 0,
                               drop_p(c, c.proj_drop, sv.d_act), c.seed, sv.d_act, c.stv);
  } else {
    linear_fwd(c, sv.h2, n, fc1, nullptr, sv.u, 0.f, 0);
    pointwise(c, sv.u, nullptr, sv.g, (int64_t)n * fc1.out, 0, c.proj_drop, sv.d_act, c.bf);
  }
  void* x3 = act(c, n, C);
  sv.d_out = drop_id(c, c.proj_drop);
  linear_fwd(c, sv.g, n, fc2, x2, x3, c.proj_drop, sv.d_out);
  return x3;
}
// D holds d(x3) on entry and d(x2) on exit
// dyd_in: mask_{d_out}(D) already produced by the LayerNorm backward that wrote D (or null); next_masked / next_did: where to leave
// mask_{next_did}(D) for the projection that follows this MLP in the backward order
static void mlp_bwd(Ctx& c, void* D, int n, const Norm& n2, const Lin& fc1, const Lin& fc2, const MlpSave& sv, const void* dyd_in = nullptr,
                    void* next_masked = nullptr, uint32_t next_did = 0) {
  int C = n2.C;
  const void* dyd = dyd_in ? dyd_in : drop_grad(c, D, (int64_t)n * C, c.proj_drop, sv.d_out, c.bf);
  void* dg = act(c, n, fc1.out);
  if (c.bf) {
    // du = drop'(gelu'(u)) * (dyd @ W2): the activation backward rides in the data-gradient GEMM's epilogue
    fork_side(c);
    if (Gf(c, fc2.w) || Gf(c, fc2.b)) CALL(lotus_linear_wgrad_bias_bf16, dyd, sv.g, Gf(c, fc2.w), Gf(c, fc2.b), n, fc2.out, fc2.in, gstream(c));   // weight + bias gradient (side stream)
    CALL(lotus_linear_dgrad_act_bias_bf16, dyd, c.WT[fc2.w], sv.u, dg, nullptr, n, fc2.out, fc2.in, 1, drop_p(c, c.proj_drop, sv.d_act),
                                          c.seed, sv.d_act, c.stv);
  } else {
    linear_bwd(c, dyd, sv.g, n, fc2, dg, nullptr, false);
    pointwise(c, sv.u, dg, dg, (int64_t)n * fc1.out, 1, c.proj_drop, sv.d_act, false);   // du in place
  }
  void* dh2 = act(c, n, C);
  linear_bwd(c, dg, sv.h2, n, fc1, dh2, nullptr, c.bf);
  if (dyd == D) join_side(c);                                                            // fc2's weight gradient reads D
  ln_bwd(c, dh2, sv.x2, D, n2, sv.mean, sv.rstd, n, 1, next_masked, next_did);           // D += LN'(dh2)
  // temporaries are released by the caller (after its join with the weight-gradient stream)
}

// ------------------------------------------------------------------------------------------------ Block
struct BlockSave {
  const void* conv_in; void *x0, *c0, *c1, *x1, *h1, *qkv, *qh, *kh, *o; float *cpe_mean, *cpe_rstd, *n1_mean, *n1_rstd, *qk_stats, *lse;
  uint32_t d_attn, d_proj; MlpSave mlp; int n; bool own_conv_in, fused;
};
static void* block_fwd(Ctx& c, const BlockP& b, const StageInfo& sg, void* x0, const void* conv_in, BlockSave& sv) {
  int n = sg.n, C = b.C, H = b.H, D = C / H;
  sv.n = n; sv.x0 = x0; sv.conv_in = conv_in ? conv_in : x0; sv.own_conv_in = conv_in != nullptr;
  sv.c0 = act(c, n, C);
  // deep stages, training only: split reduction through an fp32 scratch (the partial tiles are summed in arrival order, so inference
  // keeps the single-pass form: its output stays bit-reproducible and independent of the size bucket a CUDA-graph replay runs at)
  if (c.bf && c.train && lotus_subm_conv_splits(n, C, 27 * C) > 1) {
    float* scr = f32(c, (int64_t)n * C);                           // (a few MB at these stages; not handed back: the side stream is live)
    CALL(lotus_subm_conv_fwd_splitk_bf16, sv.conv_in, c.WB[b.conv_w], Pf(c, b.conv_b), sv.c0, sg.nbr3, sg.stride, n, C, C, 27, scr, c.stv);
  } else if (c.bf) CALL(lotus_subm_conv_fwd_bf16, sv.conv_in, c.WB[b.conv_w], Pf(c, b.conv_b), sv.c0, sg.nbr3, sg.stride, n, C, C, 27, 0, c.stv);
  else CALL(lotus_subm_conv_fwd_f32, (const float*)sv.conv_in, Pf(c, b.conv_w), Pf(c, b.conv_b), (float*)sv.c0, sg.nbr3, sg.stride, n, C, C, 27, c.stv);
  sv.c1 = act(c, n, C);
  linear_fwd(c, sv.c0, n, b.cpe1, nullptr, sv.c1, 0.f, 0);
  sv.cpe_mean = f32(c, n); sv.cpe_rstd = f32(c, n);
  sv.x1 = act(c, n, C);
  ln_fwd(c, sv.c1, sv.x1, x0, b.cpe2, sv.cpe_mean, sv.cpe_rstd, n, 1e-5f);
  sv.n1_mean = f32(c, n); sv.n1_rstd = f32(c, n);
  sv.h1 = act(c, n, C);
  ln_fwd(c, sv.x1, sv.h1, nullptr, b.n1, sv.n1_mean, sv.n1_rstd, n, 1e-5f);
  sv.qkv = act(c, n, 3 * C);
  linear_fwd(c, sv.h1, n, b.qkv, nullptr, sv.qkv, 0.f, 0);
  const void *qp = sv.qkv, *kp = (char*)sv.qkv + (size_t)C * c.es;
  const void* vp = (char*)sv.qkv + (size_t)2 * C * c.es;
  int64_t ldq = 3 * C;
  float scale = 1.0f / sqrtf((float)D);
  sv.o = act(c, n, C);
  sv.lse = f32(c, (int64_t)sg.n_pad * H);
  sv.d_attn = drop_id(c, c.attn_drop);
  sv.fused = c.bf && c.K <= 128 && D % 8 == 0 && D <= 64;
  if (sv.fused) {
    // gather + per-head q/k LayerNorm + attention + scatter in one tcgen05 kernel (attention_fused.cu)
    CALL(lotus_patch_attn_fused_fwd_bf16, qp, kp, vp, ldq, ldq, ldq, c.qk_norm ? Pf(c, b.qn.w) : nullptr, c.qk_norm ? Pf(c, b.qn.b) : nullptr,
         c.qk_norm ? Pf(c, b.kn.w) : nullptr, c.qk_norm ? Pf(c, b.kn.b) : nullptr, sg.gather_idx, sg.slot_of_point, sg.cu_seqlens, sg.n_patches,
         c.K, H, D, scale, nullptr, nullptr, 0, sv.o, C, sv.lse, drop_p(c, c.attn_drop, sv.d_attn), c.seed, sv.d_attn, c.stv);
  } else {
    if (c.qk_norm) {
      sv.qh = act(c, n, C); sv.kh = act(c, n, C);
      sv.qk_stats = f32(c, (int64_t)4 * n * H);
      head_ln_fwd(c, sv.qkv, 3 * C, sv.qh, b.qn, sv.qk_stats, sv.qk_stats + (int64_t)n * H, n, H);
      head_ln_fwd(c, (char*)sv.qkv + (size_t)C * c.es, 3 * C, sv.kh, b.kn, sv.qk_stats + (int64_t)2 * n * H, sv.qk_stats + (int64_t)3 * n * H, n, H);
      qp = sv.qh; kp = sv.kh; ldq = C;
    }
    if (c.bf) CALL(lotus_patch_attn_fwd_bf16, qp, kp, vp, ldq, ldq, 3 * C, sg.gather_idx, sg.slot_of_point, sg.cu_seqlens, sg.n_patches, c.K, H, D, scale, nullptr, nullptr, 0, sv.o, C, sv.lse, drop_p(c, c.attn_drop, sv.d_attn), c.seed, sv.d_attn, c.stv);
    else CALL(lotus_patch_attn_fwd_f32, (const float*)qp, (const float*)kp, (const float*)vp, ldq, ldq, 3 * C, sg.gather_idx, sg.slot_of_point, sg.cu_seqlens, sg.n_patches, c.K, H, D, scale, nullptr, nullptr, 0, (float*)sv.o, C, sv.lse, drop_p(c, c.attn_drop, sv.d_attn), c.seed, sv.d_attn, c.stv);
  }
  void* x2 = act(c, n, C);
  sv.d_proj = drop_id(c, c.proj_drop);
  linear_fwd(c, sv.o, n, b.proj, sv.x1, x2, c.proj_drop, sv.d_proj);
  return mlp_fwd(c, x2, n, b.n2, b.fc1, b.fc2, sv.mlp);
}
// D: d(x3) on entry -> d(x0) on exit (residual stream, in place).  If the conv read a different tensor (decoder: the skip
// branch), its gradient is returned in *d_conv_in = D_after + conv dgrad (new arena buffer, NOT released here).
static void block_bwd(Ctx& c, const BlockP& b, const StageInfo& sg, void* D, const BlockSave& sv, void** d_conv_in, const void* mlp_dyd_in = nullptr) {
  int n = sv.n, C = b.C, H = b.H, Dh = C / H;
  void* dconv = nullptr;
  if (sv.own_conv_in) dconv = act(c, n, C);                                              // survives the temporaries below
  size_t mark = c.ar.off;
  void* pm = (c.bf && sv.d_proj) ? act(c, n, C) : nullptr;                               // mask_{d_proj}(D), written by the MLP's LayerNorm backward
  mlp_bwd(c, D, n, b.n2, b.fc1, b.fc2, sv.mlp, mlp_dyd_in, pm, sv.d_proj);               // D = d(x2)
  // attention branch
  const void* dyd = pm ? pm : drop_grad(c, D, (int64_t)n * C, c.proj_drop, sv.d_proj, c.bf);
  void* dout = act(c, n, C);
  linear_bwd(c, dyd, sv.o, n, b.proj, dout, nullptr, c.bf);
  void* dqkv = act(c, n, 3 * C);
  float scale = 1.0f / sqrtf((float)Dh);
  const void* vp = (char*)sv.qkv + (size_t)2 * C * c.es;
  if (sv.fused) {
    // every row of dqkv is written by the kernel (bf16, raw-space gradients: the per-head LayerNorm backward runs in its epilogue)
    void* alias = c.ar.alloc((size_t)sg.n_pad * 2 * C * 2);
    CALL(lotus_patch_attn_fused_bwd_bf16, sv.qkv, (char*)sv.qkv + (size_t)C * 2, vp, 3 * C, 3 * C, 3 * C, c.qk_norm ? Pf(c, b.qn.w) : nullptr,
         c.qk_norm ? Pf(c, b.qn.b) : nullptr, c.qk_norm ? Pf(c, b.kn.w) : nullptr, c.qk_norm ? Pf(c, b.kn.b) : nullptr, sg.gather_idx, sg.slot_of_point,
         sg.cu_seqlens, sg.n_patches, sg.n_pad, c.K, H, Dh, scale, nullptr, nullptr, 0, nullptr, sv.o, dout, C, sv.lse, dqkv, (char*)dqkv + (size_t)C * 2,
         (char*)dqkv + (size_t)2 * C * 2, 3 * C, 3 * C, 3 * C, Gf(c, b.qn.w), Gf(c, b.qn.b), Gf(c, b.kn.w), Gf(c, b.kn.b), alias,
         drop_p(c, c.attn_drop, sv.d_attn), c.seed, sv.d_attn, c.stv);
  } else if (c.bf || c.qk_norm) {
    float* dqh = f32(c, (int64_t)n * C); float* dkh = f32(c, (int64_t)n * C);
    float* dvh = c.bf ? f32(c, (int64_t)n * C) : (float*)dqkv + 2 * C;
    int64_t lddv = c.bf ? C : 3 * C;
    zero(c, dqh, (size_t)n * C * 4); zero(c, dkh, (size_t)n * C * 4);
    if (c.bf) zero(c, dvh, (size_t)n * C * 4); else zero(c, dqkv, (size_t)n * 3 * C * 4);
    const void *qp = c.qk_norm ? sv.qh : sv.qkv, *kp = c.qk_norm ? sv.kh : (void*)((char*)sv.qkv + (size_t)C * c.es);
    int64_t ldq = c.qk_norm ? C : 3 * C;
    if (c.bf) CALL(lotus_patch_attn_bwd_bf16, qp, kp, vp, ldq, ldq, 3 * C, sg.gather_idx, sg.slot_of_point, sg.cu_seqlens, sg.n_patches, c.K, H, Dh, scale, nullptr, nullptr, 0, nullptr, sv.o, dout, C, sv.lse, dqh, dkh, dvh, C, C, lddv, drop_p(c, c.attn_drop, sv.d_attn), c.seed, sv.d_attn, c.stv);
    else CALL(lotus_patch_attn_bwd_f32, (const float*)qp, (const float*)kp, (const float*)vp, ldq, ldq, 3 * C, sg.gather_idx, sg.slot_of_point, sg.cu_seqlens, sg.n_patches, c.K, H, Dh, scale, nullptr, nullptr, 0, nullptr, (const float*)sv.o, (const float*)dout, C, sv.lse, dqh, dkh, dvh, C, C, lddv, drop_p(c, c.attn_drop, sv.d_attn), c.seed, sv.d_attn, c.stv);
    if (c.qk_norm) {
      int64_t nh = (int64_t)n * H;
      head_ln_bwd(c, dqh, sv.qkv, 3 * C, dqkv, b.qn, sv.qk_stats, sv.qk_stats + nh, n, H);
      head_ln_bwd(c, dkh, (char*)sv.qkv + (size_t)C * c.es, 3 * C, (char*)dqkv + (size_t)C * c.es, b.kn, sv.qk_stats + 2 * nh, sv.qk_stats + 3 * nh, n, H);
    } else {
      CALL(lotus_convert_rows_f32_to_bf16, dqh, C, dqkv, 3 * C, n, C, c.stv);
      CALL(lotus_convert_rows_f32_to_bf16, dkh, C, (char*)dqkv + (size_t)C * 2, 3 * C, n, C, c.stv);
    }
    if (c.bf) CALL(lotus_convert_rows_f32_to_bf16, dvh, C, (char*)dqkv + (size_t)2 * C * 2, 3 * C, n, C, c.stv);
  } else {
    zero(c, dqkv, (size_t)n * 3 * C * 4);
    float* dq = (float*)dqkv;
    CALL(lotus_patch_attn_bwd_f32, (const float*)sv.qkv, (const float*)sv.qkv + C, (const float*)vp, 3 * C, 3 * C, 3 * C, sg.gather_idx, sg.slot_of_point, sg.cu_seqlens, sg.n_patches, c.K, H, Dh, scale, nullptr, nullptr, 0, nullptr, (const float*)sv.o, (const float*)dout, C, sv.lse, dq, dq + C, dq + 2 * C, 3 * C, 3 * C, 3 * C, drop_p(c, c.attn_drop, sv.d_attn), c.seed, sv.d_attn, c.stv);
  }
  void* dh1 = act(c, n, C);
  linear_bwd(c, dqkv, sv.h1, n, b.qkv, dh1, nullptr, c.bf);
  if (dyd == D) join_side(c);                                                            // proj's weight gradient reads D
  ln_bwd(c, dh1, sv.x1, D, b.n1, sv.n1_mean, sv.n1_rstd, n, 1);                          // D = d(x1)
  // CPE branch: x1 = x0 + LN(c1)
  void* dc1 = act(c, n, C);
  ln_bwd(c, D, sv.c1, dc1, b.cpe2, sv.cpe_mean, sv.cpe_rstd, n, 0);
  void* dc0 = act(c, n, C);
  linear_bwd(c, dc1, sv.c0, n, b.cpe1, dc0, nullptr, c.bf);
  // conv: dW, db, and the input gradient (into the residual stream, or into the separate conv-input gradient)
  void* dxo = sv.own_conv_in ? dconv : D;
  fork_side(c);
  void* gs = gstream(c);
  if (c.bf) {
    if (Gf(c, b.conv_w)) CALL(lotus_subm_conv_wgrad_pairs_bf16, dc0, sv.conv_in, Gf(c, b.conv_w), sg.pairs_out, sg.pairs_in, sg.pair_cnt, sg.stride, n, C, C, 27, gs);
    if (Gf(c, b.conv_b)) CALL(lotus_colsum_bf16, dc0, n, C, Gf(c, b.conv_b), gs);
    if (lotus_subm_conv_splits(n, C, 27 * C) > 1) {
      float* scr = f32(c, (int64_t)n * C);
      CALL(lotus_subm_conv_dgrad_splitk_bf16, dc0, c.WT[b.conv_w], D, dxo, sg.nbr3, sg.stride, n, C, C, 27, scr, c.stv);
    } else CALL(lotus_subm_conv_dgrad_bf16, dc0, c.WT[b.conv_w], D, dxo, sg.nbr3, sg.stride, n, C, C, 27, 0, c.stv);
  } else {
    if (Gf(c, b.conv_w)) CALL(lotus_subm_conv_wgrad_f32, (const float*)dc0, (const float*)sv.conv_in, Gf(c, b.conv_w), sg.nbr3, sg.stride, n, C, C, 27, gs);
    if (Gf(c, b.conv_b)) CALL(lotus_colsum_f32, (const float*)dc0, n, C, Gf(c, b.conv_b), gs);
    CALL(lotus_subm_conv_dgrad_f32, (const float*)dc0, Pf(c, b.conv_w), (const float*)D, (float*)dxo, sg.nbr3, sg.stride, n, C, C, 27, c.stv);
  }
  join_side(c);                                                                          // before the temporaries are reused
  c.ar.off = mark;
  if (d_conv_in) *d_conv_in = dconv;
}

// ------------------------------------------------------------------------------------------------ CABlock
struct CaSave { void *x0, *h1, *q, *kv, *qh, *kh, *o; float *n1_mean, *n1_rstd, *q_stats, *k_stats, *lse; uint32_t d_attn, d_proj; MlpSave mlp; int n; bool kv_ready = false; };
// keys / values of a CABlock depend on the text tokens only: all blocks' kv projections (+ per-head k LayerNorm) are issued up
// front on the side stream, off the point-feature chain (18 small launches per forward)
static void ca_kv(Ctx& c, const CaP& b, const void* ctx, CaSave& sv) {
  int C = b.C, H = b.H, Lt = c.Lt;
  sv.kv = act(c, Lt, 2 * C);
  linear_fwd(c, ctx, Lt, b.kv, nullptr, sv.kv, 0.f, 0);
  if (c.qk_norm) {
    sv.kh = act(c, Lt, C);
    sv.k_stats = f32(c, (int64_t)2 * Lt * H);
    head_ln_fwd(c, sv.kv, 2 * C, sv.kh, b.kn, sv.k_stats, sv.k_stats + (int64_t)Lt * H, Lt, H);
  }
  sv.kv_ready = true;
}
static void* ca_fwd(Ctx& c, const CaP& b, const StageInfo& sg, void* x0, const void* ctx, CaSave& sv) {
  int n = sg.n, C = b.C, H = b.H, D = C / H;
  sv.n = n; sv.x0 = x0;
  sv.n1_mean = f32(c, n); sv.n1_rstd = f32(c, n);
  sv.h1 = act(c, n, C);
  ln_fwd(c, x0, sv.h1, nullptr, b.n1, sv.n1_mean, sv.n1_rstd, n, 1e-5f);
  sv.q = act(c, n, C);
  linear_fwd(c, sv.h1, n, b.q, nullptr, sv.q, 0.f, 0);
  if (!sv.kv_ready) ca_kv(c, b, ctx, sv);
  const void *qp = sv.q, *kp = sv.kv; int64_t ldk = 2 * C;
  if (c.qk_norm) {
    sv.qh = act(c, n, C);
    sv.q_stats = f32(c, (int64_t)2 * n * H);
    head_ln_fwd(c, sv.q, C, sv.qh, b.qn, sv.q_stats, sv.q_stats + (int64_t)n * H, n, H);
    qp = sv.qh; kp = sv.kh; ldk = C;
  }
  join_side(c);                                       // the up-front kv projections (no-op after the first CABlock)
  sv.o = act(c, n, C);
  sv.lse = f32(c, (int64_t)n * H);
  sv.d_attn = drop_id(c, c.attn_drop);
  const void* vp = (char*)sv.kv + (size_t)C * c.es;
  float scale = 1.0f / sqrtf((float)D);
  if (c.bf) CALL(lotus_cross_attn_fwd_bf16, qp, C, kp, vp, ldk, 2 * C, sg.cloud_off, c.tok_off, c.B, sg.max_cloud, c.max_tok, H, D, scale, sv.o, C, sv.lse, drop_p(c, c.attn_drop, sv.d_attn), c.seed, sv.d_attn, c.stv);
  else CALL(lotus_cross_attn_fwd_f32, (const float*)qp, C, (const float*)kp, (const float*)vp, ldk, 2 * C, sg.cloud_off, c.tok_off, c.B, sg.max_cloud, c.max_tok, H, D, scale, (float*)sv.o, C, sv.lse, drop_p(c, c.attn_drop, sv.d_attn), c.seed, sv.d_attn, c.stv);
  void* x2 = act(c, n, C);
  sv.d_proj = drop_id(c, c.proj_drop);
  linear_fwd(c, sv.o, n, b.proj, x0, x2, c.proj_drop, sv.d_proj);
  return mlp_fwd(c, x2, n, b.n2, b.fc1, b.fc2, sv.mlp);
}
// D: d(out) -> d(x0) in place; dctx (Lt x ctxC, activation dtype) accumulates (first use flagged by *dctx_init)
static void ca_bwd(Ctx& c, const CaP& b, const StageInfo& sg, void* D, const void* ctx, void* dctx, bool* dctx_init, const CaSave& sv,
                   void* next_masked = nullptr, uint32_t next_did = 0) {
  int n = sv.n, C = b.C, H = b.H, Dh = C / H, Lt = c.Lt;
  size_t mark = c.ar.off;
  void* pm = (c.bf && sv.d_proj) ? act(c, n, C) : nullptr;
  mlp_bwd(c, D, n, b.n2, b.fc1, b.fc2, sv.mlp, nullptr, pm, sv.d_proj);
  const void* dyd = pm ? pm : drop_grad(c, D, (int64_t)n * C, c.proj_drop, sv.d_proj, c.bf);
  void* dout = act(c, n, C);
  linear_bwd(c, dyd, sv.o, n, b.proj, dout, nullptr, c.bf);
  void* dq = act(c, n, C);
  void* dkv = act(c, Lt, 2 * C);
  float scale = 1.0f / sqrtf((float)Dh);
  const void* vp = (char*)sv.kv + (size_t)C * c.es;
  if (c.bf || c.qk_norm) {
    float* dqh = f32(c, (int64_t)n * C); float* dkh = f32(c, (int64_t)Lt * C);
    float* dvh = c.bf ? f32(c, (int64_t)Lt * C) : (float*)dkv + C;
    int64_t lddv = c.bf ? C : 2 * C;
    zero(c, dkh, (size_t)Lt * C * 4);
    if (c.bf) zero(c, dvh, (size_t)Lt * C * 4); else zero(c, dkv, (size_t)Lt * 2 * C * 4);
    const void *qp = c.qk_norm ? sv.qh : sv.q, *kp = c.qk_norm ? sv.kh : sv.kv;
    int64_t ldk = c.qk_norm ? C : 2 * C;
    if (c.bf) CALL(lotus_cross_attn_bwd_bf16, qp, C, kp, vp, ldk, 2 * C, sg.cloud_off, c.tok_off, c.B, sg.max_cloud, c.max_tok, H, Dh, scale, sv.o, dout, C, sv.lse, dqh, C, dkh, dvh, C, lddv, drop_p(c, c.attn_drop, sv.d_attn), c.seed, sv.d_attn, c.stv);
    else CALL(lotus_cross_attn_bwd_f32, (const float*)qp, C, (const float*)kp, (const float*)vp, ldk, 2 * C, sg.cloud_off, c.tok_off, c.B, sg.max_cloud, c.max_tok, H, Dh, scale, (const float*)sv.o, (const float*)dout, C, sv.lse, dqh, C, dkh, dvh, C, lddv, drop_p(c, c.attn_drop, sv.d_attn), c.seed, sv.d_attn, c.stv);
    if (c.qk_norm) {
      head_ln_bwd(c, dqh, sv.q, C, dq, b.qn, sv.q_stats, sv.q_stats + (int64_t)n * H, n, H);
      head_ln_bwd(c, dkh, sv.kv, 2 * C, dkv, b.kn, sv.k_stats, sv.k_stats + (int64_t)Lt * H, Lt, H);
    } else {
      CALL(lotus_convert_rows_f32_to_bf16, dqh, C, dq, C, n, C, c.stv);
      CALL(lotus_convert_rows_f32_to_bf16, dkh, C, dkv, 2 * C, Lt, C, c.stv);
    }
    if (c.bf) CALL(lotus_convert_rows_f32_to_bf16, dvh, C, (char*)dkv + (size_t)C * 2, 2 * C, Lt, C, c.stv);
  } else {
    zero(c, dkv, (size_t)Lt * 2 * C * 4);
    CALL(lotus_cross_attn_bwd_f32, (const float*)sv.q, C, (const float*)sv.kv, (const float*)vp, 2 * C, 2 * C, sg.cloud_off, c.tok_off, c.B, sg.max_cloud, c.max_tok, H, Dh, scale, (const float*)sv.o, (const float*)dout, C, sv.lse, (float*)dq, C, (float*)dkv, (float*)dkv + C, 2 * C, 2 * C, drop_p(c, c.attn_drop, sv.d_attn), c.seed, sv.d_attn, c.stv);
  }
  linear_bwd(c, dkv, ctx, Lt, b.kv, dctx, *dctx_init ? dctx : nullptr, c.bf);           // dctx (+)= dkv @ Wkv
  *dctx_init = true;
  void* dh1 = act(c, n, C);
  linear_bwd(c, dq, sv.h1, n, b.q, dh1, nullptr, c.bf);
  if (dyd == D) join_side(c);
  ln_bwd(c, dh1, sv.x0, D, b.n1, sv.n1_mean, sv.n1_rstd, n, 1, next_masked, next_did);   // (+ the mask for the block MLP that follows)
  join_side(c);
  c.ar.off = mark;
}

// ------------------------------------------------------------------------------------------------ model description
struct Model {
  int S; std::vector<int> encC, encH, decC, decH; int inC, txtC, ctxC, pos_bins, rot_bins;
  int stem_w; BNorm stem_bn;
  std::vector<BlockP> encB, decB; std::vector<CaP> encCA, decCA; std::vector<PoolP> pool; std::vector<UnpoolP> unpool;
  Lin txt_fc, hm0, hm3, am0, am3; int n_params, n_bn;
};
struct Cursor { int i = 0; int bn = 0; };
static Lin take_lin(Cursor& k, int out, int in) { Lin l{k.i, k.i + 1, out, in}; k.i += 2; return l; }
static Norm take_norm(Cursor& k, int C) { Norm n{k.i, k.i + 1, C}; k.i += 2; return n; }
static BNorm take_bn(Cursor& k, int C) { BNorm b{k.i, k.i + 1, C, k.bn++}; k.i += 2; return b; }
static BlockP take_block(Cursor& k, int C, int H) {
  BlockP b; b.C = C; b.H = H;
  b.conv_w = k.i; b.conv_b = k.i + 1; k.i += 2;
  b.cpe1 = take_lin(k, C, C); b.cpe2 = take_norm(k, C); b.n1 = take_norm(k, C);
  b.qkv = take_lin(k, 3 * C, C); b.proj = take_lin(k, C, C);
  b.qn = take_norm(k, C / H); b.kn = take_norm(k, C / H);
  b.n2 = take_norm(k, C); b.fc1 = take_lin(k, 4 * C, C); b.fc2 = take_lin(k, C, 4 * C);
  return b;
}
static CaP take_ca(Cursor& k, int C, int H, int ctxC) {
  CaP b; b.C = C; b.H = H;
  b.n1 = take_norm(k, C); b.q = take_lin(k, C, C); b.kv = take_lin(k, 2 * C, ctxC); b.proj = take_lin(k, C, C);
  b.qn = take_norm(k, C / H); b.kn = take_norm(k, C / H);
  b.n2 = take_norm(k, C); b.fc1 = take_lin(k, 4 * C, C); b.fc2 = take_lin(k, C, 4 * C);
  return b;
}
// parameter order == named_parameters() of SimplePolicyPTV3CA
static Model build_model(const int64_t* cfg) {
  Model m; int p = 0;
  m.S = (int)cfg[p++]; m.inC = (int)cfg[p++]; m.txtC = (int)cfg[p++]; m.ctxC = (int)cfg[p++]; m.pos_bins = (int)cfg[p++]; m.rot_bins = (int)cfg[p++];
  for (int s = 0; s < m.S; ++s) m.encC.push_back((int)cfg[p++]);
  for (int s = 0; s < m.S; ++s) m.encH.push_back((int)cfg[p++]);
  for (int s = 0; s < m.S - 1; ++s) m.decC.push_back((int)cfg[p++]);
  for (int s = 0; s < m.S - 1; ++s) m.decH.push_back((int)cfg[p++]);
  Cursor k;
  m.stem_w = k.i++; m.stem_bn = take_bn(k, m.encC[0]);
  m.encB.resize(m.S); m.encCA.resize(m.S); m.pool.resize(m.S);
  for (int s = 0; s < m.S; ++s) {
    if (s > 0) { m.pool[s].proj = take_lin(k, m.encC[s], m.encC[s - 1]); m.pool[s].bn = take_bn(k, m.encC[s]); }
    m.encB[s] = take_block(k, m.encC[s], m.encH[s]);
    m.encCA[s] = take_ca(k, m.encC[s], m.encH[s], m.ctxC);
  }
  m.decB.resize(m.S - 1); m.decCA.resize(m.S - 1); m.unpool.resize(m.S - 1);
  for (int s = m.S - 2; s >= 0; --s) {
    int cin = (s == m.S - 2) ? m.encC[m.S - 1] : m.decC[s + 1];
    m.unpool[s].proj = take_lin(k, m.decC[s], cin); m.unpool[s].bn = take_bn(k, m.decC[s]);
    m.unpool[s].skip = take_lin(k, m.decC[s], m.encC[s]); m.unpool[s].bn_skip = take_bn(k, m.decC[s]);
    m.decB[s] = take_block(k, m.decC[s], m.decH[s]);
    m.decCA[s] = take_ca(k, m.decC[s], m.decH[s], m.ctxC);
  }
  m.txt_fc = take_lin(k, m.ctxC, m.txtC);
  int hid = m.decC[0];
  m.hm0 = take_lin(k, hid, hid); m.hm3 = take_lin(k, 3 * 2 * m.pos_bins, hid);
  m.am0 = take_lin(k, hid, hid); m.am3 = take_lin(k, m.rot_bins * 3 + 1, hid);
  m.n_params = k.i; m.n_bn = k.bn;
  return m;
}

// weights to stage this step (flushed as ONE batched cast/transpose launch)
struct StageList { std::vector<const void*> w; std::vector<void*> wb, wt; std::vector<int> A, D, B; };
static StageList g_stage;
static void stage_weight(Ctx& c, int idx, int A, int D, int Bdim) {
  c.WB[idx] = c.ar.alloc((size_t)A * D * Bdim * 2);
  c.WT[idx] = c.ar.alloc((size_t)A * D * Bdim * 2);
  g_stage.w.push_back(Pf(c, idx)); g_stage.wb.push_back(c.WB[idx]); g_stage.wt.push_back(c.WT[idx]);
  g_stage.A.push_back(A); g_stage.D.push_back(D); g_stage.B.push_back(Bdim);
}
static void stage_flush(Ctx& c, bool skip) {
  if (!skip) CALL(lotus_cast_transpose_batch_bf16, g_stage.w.data(), g_stage.wb.data(), g_stage.wt.data(), g_stage.A.data(), g_stage.D.data(),
                                       g_stage.B.data(), (int)g_stage.w.size(), c.stv);
  g_stage = StageList();
}
static void stage_lin(Ctx& c, const Lin& l) { stage_weight(c, l.w, l.out, 1, l.in); }
static void stage_block(Ctx& c, const BlockP& b) {
  stage_weight(c, b.conv_w, b.C, 27, b.C);
  stage_lin(c, b.cpe1); stage_lin(c, b.qkv); stage_lin(c, b.proj); stage_lin(c, b.fc1); stage_lin(c, b.fc2);
}
static void stage_ca(Ctx& c, const CaP& b) { stage_lin(c, b.q); stage_lin(c, b.kv); stage_lin(c, b.proj); stage_lin(c, b.fc1); stage_lin(c, b.fc2); }

}  // namespace eng

using namespace eng;

// cfg (int64): S, in_channels, txt_dim, ctx_channels, pos_bins, rot_bins, enc_channels[S], enc_heads[S], dec_channels[S-1],
//              dec_heads[S-1]
// run (int64): B, Lt, max_tokens, patch_size K, qk_norm, train, bf16, compute_loss, use_gt_pos, dry_run, weights_staged
//   (weights_staged = 1: the bf16 weight copies at the head of `arena` are still valid from the previous call -- the
//    caller guarantees same arena base and unchanged parameters; used by repeated inference calls)
// stage_sizes (int64, 5 per stage): n, n_pad, n_patches, max_cloud, stride ; stage_ptrs (12 per stage): nbr3, pairs_out, pairs_in,
//   pair_cnt, gather_idx, slot_of_point, cu_seqlens, cloud_off, cluster, seg_ptr, member, nbr5
// fl (float): attn_drop, proj_drop, head_drop, pos_weight, rot_weight, euler_resolution ; dbl: pos_bin_size
// slice_events (host array of 3 cudaEvent_t, or null): recorded on `stream` during the backward when the parameter gradients of
//   (0) head + decoder, (1) the deepest encoder stage, (2) the stage above it are complete (weight-gradient side stream joined), so the
//   caller can start the gradient all-reduce of those slices while the rest of the backward runs (train/utils/distributed.py:196-212:
//   DDP overlaps its bucketed all-reduces with the backward the same way)
// Returns 0 and writes losses[4] (device), final actions (B,8 double, device); *arena_peak receives the bytes used
// (with dry_run = 1 nothing is launched: call once to size the arena).
LOTUS_API int lotus_policy_step(const int64_t* cfg, const int64_t* run, const int64_t* stage_sizes, void* const* stage_ptrs,
                                void* const* params, void* const* grads, void* const* bn_stats, const float* feat,
                                const float* txt, const int32_t* tok_off, const float* gt_actions, const float* pos_prob,
                                const float* fl, double pos_bin_size, uint64_t seed, void* arena, int64_t arena_bytes,
                                int64_t* arena_peak, float* losses, double* final_actions, void* const* slice_events, void* stream_) {
  Model m = build_model(cfg);
  Ctx c;
  if (run[9] == 0) g_launcher_calls = 0;
  c.st = (cudaStream_t)stream_; c.stv = stream_;
  c.B = (int)run[0]; c.Lt = (int)run[1]; c.max_tok = (int)run[2]; c.K = (int)run[3]; c.qk_norm = (int)run[4];
  c.train = run[5] != 0; c.bf = run[6] != 0; bool compute_loss = run[7] != 0, use_gt_pos = run[8] != 0; c.dry = run[9] != 0;
  c.es = c.bf ? 2 : 4; c.rc = 0; c.seed = seed; c.drop_next = 1;
  c.side = nullptr; c.sidev = nullptr; c.side_busy = false;
  if (!c.dry) { g_prof.clear(); g_prof_ev_next = 0; }
  if (!c.dry && !g_prof_on && getenv("LOTUS_NO_SIDE_STREAM") == nullptr) {
    if (!g_side) {
      LOTUS_CUDA_CALL(cudaStreamCreateWithFlags(&g_side, cudaStreamNonBlocking));
      LOTUS_CUDA_CALL(cudaEventCreateWithFlags(&g_ev_fork, cudaEventDisableTiming));
      LOTUS_CUDA_CALL(cudaEventCreateWithFlags(&g_ev_join, cudaEventDisableTiming));
    }
    c.side = g_side; c.sidev = (void*)g_side;
  }
  c.attn_drop = fl[0]; c.proj_drop = fl[1]; c.head_drop = fl[2];
  float pos_w = fl[3], rot_w = fl[4], euler_res = fl[5];
  c.P = params; c.G = grads; c.BN = bn_stats; c.tok_off = tok_off; c.S = m.S; c.ctxC = m.ctxC;
  c.ar.base = (char*)arena; c.ar.off = 0; c.ar.cap = (size_t)arena_bytes; c.ar.peak = 0; c.ar.dry = c.dry; c.ar.overflow = false;
  c.WB.assign(m.n_params, nullptr); c.WT.assign(m.n_params, nullptr);
  const bool need_grad = grads != nullptr && compute_loss;
  if (!need_grad) c.G = nullptr;
  int S = m.S, B = c.B;
  std::vector<StageInfo> sg(S);
  for (int s = 0; s < S; ++s) {
    const int64_t* z = stage_sizes + 5 * s; void* const* q = stage_ptrs + 12 * s;
    sg[s].n = (int)z[0]; sg[s].n_pad = (int)z[1]; sg[s].n_patches = (int)z[2]; sg[s].max_cloud = (int)z[3]; sg[s].stride = z[4];
    sg[s].nbr3 = (const int32_t*)q[0]; sg[s].pairs_out = (const int32_t*)q[1]; sg[s].pairs_in = (const int32_t*)q[2];
    sg[s].pair_cnt = (const int32_t*)q[3]; sg[s].gather_idx = (const int32_t*)q[4]; sg[s].slot_of_point = (const int32_t*)q[5];
    sg[s].cu_seqlens = (const int32_t*)q[6]; sg[s].cloud_off = (const int32_t*)q[7]; sg[s].cluster = (const int32_t*)q[8];
    sg[s].seg_ptr = (const int32_t*)q[9]; sg[s].member = (const int32_t*)q[10]; sg[s].nbr5 = (const int32_t*)q[11];
  }
  if (c.bf) {
    for (int s = 0; s < S; ++s) if ((m.encC[s] % 64) || (s < S - 1 && (m.decC[s] % 64))) {
      lotus_set_error("lotus_policy_step: bf16 mode needs channel widths that are multiples of 64");
      return 2;
    }
    for (int s = 0; s < S; ++s) {
      if (s > 0) stage_lin(c, m.pool[s].proj);
      stage_block(c, m.encB[s]); stage_ca(c, m.encCA[s]);
      if (s < S - 1) { stage_lin(c, m.unpool[s].proj); stage_lin(c, m.unpool[s].skip); stage_block(c, m.decB[s]); stage_ca(c, m.decCA[s]); }
    }
    stage_lin(c, m.txt_fc); stage_lin(c, m.hm0); stage_lin(c, m.am0);
    stage_flush(c, run[10] != 0);      // run[10]: the arena still holds this model's staged weights (inference)
  }
  const int N0 = sg[0].n, Lt = c.Lt;
  // ---------------- context tokens
  const void* txt_in = txt;
  if (c.bf) { void* t = c.ar.alloc((size_t)Lt * m.txtC * 2); CALL(lotus_convert, txt, t, (int64_t)Lt * m.txtC, 0, c.stv); txt_in = t; }
  void* ctx = act(c, Lt, m.ctxC);
  linear_fwd(c, txt_in, Lt, m.txt_fc, nullptr, ctx, 0.f, 0);
  std::vector<BlockSave> eB(S), dB(S - 1); std::vector<CaSave> eC(S), dC(S - 1);
  if (c.side) {
    fork_side(c);                                     // ctx is complete on the main stream
    cudaStream_t st_main = c.st; void* stv_main = c.stv;
    c.st = c.side; c.stv = c.sidev;
    for (int s = 0; s < S; ++s) ca_kv(c, m.encCA[s], ctx, eC[s]);
    for (int s = S - 2; s >= 0; --s) ca_kv(c, m.decCA[s], ctx, dC[s]);
    c.st = st_main; c.stv = stv_main;
  }
  // ---------------- stem (fp32 SIMT conv, narrow input) + BN + GELU
  float* stem_out = f32(c, (int64_t)N0 * m.encC[0]);
  CALL(lotus_subm_conv_fwd_f32, feat, Pf(c, m.stem_w), nullptr, stem_out, sg[0].nbr5, sg[0].stride, N0, m.inC, m.encC[0], 125, c.stv);
  BnSave stem_sv;
  void* x = act(c, N0, m.encC[0]);
  bn_fwd(c, stem_out, true, x, N0, m.stem_bn, stem_sv);
  // ---------------- encoder
  std::vector<BnSave> poolBn(S), upBn(S - 1), upBnSkip(S - 1);
  std::vector<void*> skip(S, nullptr), pool_proj(S, nullptr), pooled(S, nullptr), up_lin(S - 1, nullptr), up_skip_lin(S - 1, nullptr),
      up_a(S - 1, nullptr), up_b(S - 1, nullptr), up_in(S - 1, nullptr);
  std::vector<int32_t*> pool_arg(S, nullptr);
  for (int s = 0; s < S; ++s) {
    if (s > 0) {
      skip[s - 1] = x;
      int np = sg[s - 1].n, n = sg[s].n, C = m.encC[s];
      pool_proj[s] = act(c, np, C);
      linear_fwd(c, x, np, m.pool[s].proj, nullptr, pool_proj[s], 0.f, 0);
      pooled[s] = act(c, n, C);
      pool_arg[s] = (int32_t*)c.ar.alloc((size_t)n * C * 4);
      if (c.bf) CALL(lotus_segment_max_fwd_bf16, pool_proj[s], C, sg[s].seg_ptr, sg[s].member, n, pooled[s], pool_arg[s], c.stv);
      else CALL(lotus_segment_max_fwd_f32, (const float*)pool_proj[s], C, sg[s].seg_ptr, sg[s].member, n, (float*)pooled[s], pool_arg[s], c.stv);
      x = act(c, n, C);
      bn_fwd(c, pooled[s], false, x, n, m.pool[s].bn, poolBn[s]);
    }
    x = block_fwd(c, m.encB[s], sg[s], x, nullptr, eB[s]);
    x = ca_fwd(c, m.encCA[s], sg[s], x, ctx, eC[s]);
  }
  // ---------------- decoder
  for (int s = S - 2; s >= 0; --s) {
    int nc = sg[s + 1].n, n = sg[s].n, C = m.decC[s];
    up_in[s] = x;
    up_lin[s] = act(c, nc, C);
    linear_fwd(c, x, nc, m.unpool[s].proj, nullptr, up_lin[s], 0.f, 0);
    up_a[s] = act(c, nc, C);
    bn_fwd(c, up_lin[s], false, up_a[s], nc, m.unpool[s].bn, upBn[s]);
    up_skip_lin[s] = act(c, n, C);
    linear_fwd(c, skip[s], n, m.unpool[s].skip, nullptr, up_skip_lin[s], 0.f, 0);
    up_b[s] = act(c, n, C);
    bn_fwd(c, up_skip_lin[s], false, up_b[s], n, m.unpool[s].bn_skip, upBnSkip[s]);
    void* xu = act(c, n, C);
    if (c.bf) CALL(lotus_unpool_add_bf16, up_b[s], up_a[s], sg[s + 1].cluster, n, C, xu, c.stv);
    else CALL(lotus_unpool_add_f32, (const float*)up_b[s], (const float*)up_a[s], sg[s + 1].cluster, n, C, (float*)xu, c.stv);
    x = block_fwd(c, m.decB[s], sg[s], xu, up_b[s], dB[s]);          // the CPE conv reads the skip branch only (model.py:824 vs 235-238)
    x = ca_fwd(c, m.decCA[s], sg[s], x, ctx, dC[s]);
  }
  // ---------------- head
  int hid = m.decC[0], NB = 2 * m.pos_bins, hw = 3 * NB, aw = m.rot_bins * 3 + 1;
  float* h = f32(c, (int64_t)N0 * hid);
  linear_fwd(c, x, N0, m.hm0, nullptr, h, 0.f, 0, true);
  float* hl = f32(c, (int64_t)N0 * hid);
  uint32_t d_h = drop_id(c, c.head_drop);
  pointwise(c, h, nullptr, hl, (int64_t)N0 * hid, 2, c.head_drop, d_h, false);
  float* heat = f32(c, (int64_t)N0 * hw);
  linear_fwd(c, hl, N0, m.hm3, nullptr, heat, 0.f, 0, true, false);
  void* pc = act(c, B, hid);
  int32_t* pc_arg = (int32_t*)c.ar.alloc((size_t)B * hid * 4);
  if (c.bf) CALL(lotus_segment_max_fwd_bf16, x, hid, sg[0].cloud_off, nullptr, B, pc, pc_arg, c.stv);
  else CALL(lotus_segment_max_fwd_f32, (const float*)x, hid, sg[0].cloud_off, nullptr, B, (float*)pc, pc_arg, c.stv);
  float* h2 = f32(c, (int64_t)B * hid);
  linear_fwd(c, pc, B, m.am0, nullptr, h2, 0.f, 0, true);
  float* hl2 = f32(c, (int64_t)B * hid);
  uint32_t d_h2 = drop_id(c, c.head_drop);
  pointwise(c, h2, nullptr, hl2, (int64_t)B * hid, 2, c.head_drop, d_h2, false);
  float* action = f32(c, (int64_t)B * aw);
  linear_fwd(c, hl2, B, m.am3, nullptr, action, 0.f, 0, true, false);
  CALL(lotus_decode_actions, heat, feat, m.inC, sg[0].cloud_off, B, NB, pos_bin_size, action, aw, m.rot_bins, euler_res,
                            use_gt_pos ? gt_actions : nullptr, 7, final_actions, c.stv);
  if (compute_loss) {
    float* dheat = need_grad ? f32(c, (int64_t)N0 * hw) : nullptr;
    float* daction = need_grad ? f32(c, (int64_t)B * aw) : nullptr;
    CALL(lotus_policy_loss_f32, heat, pos_prob, sg[0].cloud_off, B, NB, action, aw, m.rot_bins, gt_actions, 7, pos_w, rot_w, losses, dheat, daction, nullptr, c.stv);
    if (need_grad) {
      // ================= backward =================
      // head
      void* D = act(c, N0, hid);                                     // gradient of the residual stream at stage 0 (decoder side)
      {
        size_t mark = c.ar.off;
        float* dhl = f32(c, (int64_t)N0 * hid);
        linear_bwd(c, dheat, hl, N0, m.hm3, dhl, nullptr, false);
        pointwise(c, h, dhl, dhl, (int64_t)N0 * hid, 3, c.head_drop, d_h, false);       // dh (fp32) in place
        const void* dh = dhl;
        if (c.bf) { void* t = c.ar.alloc((size_t)N0 * hid * 2); CALL(lotus_convert, dhl, t, (int64_t)N0 * hid, 0, c.stv); dh = t; }
        linear_bwd(c, dh, x, N0, m.hm0, D, nullptr, c.bf);
        float* dhl2 = f32(c, (int64_t)B * hid);
        linear_bwd(c, daction, hl2, B, m.am3, dhl2, nullptr, false);
        pointwise(c, h2, dhl2, dhl2, (int64_t)B * hid, 3, c.head_drop, d_h2, false);
        const void* dh2 = dhl2;
        if (c.bf) { void* t = c.ar.alloc((size_t)B * hid * 2); CALL(lotus_convert, dhl2, t, (int64_t)B * hid, 0, c.stv); dh2 = t; }
        void* dpc = act(c, B, hid);
        linear_bwd(c, dh2, pc, B, m.am0, dpc, nullptr, c.bf);
        if (c.bf) CALL(lotus_segment_max_bwd_bf16, dpc, pc_arg, B, hid, D, 1, c.stv);
        else CALL(lotus_segment_max_bwd_f32, (const float*)dpc, pc_arg, B, hid, (float*)D, 1, c.stv);
        join_side(c);
        c.ar.off = mark;
      }
      void* dctx = act(c, Lt, m.ctxC);
      bool dctx_init = false;
      std::vector<void*> Gskip(S, nullptr);                          // gradient w.r.t. each encoder stage output coming from the decoder skip
      // decoder, shallow to deep
      for (int s = 0; s <= S - 2; ++s) {
        int nc = sg[s + 1].n, n = sg[s].n, C = m.decC[s];
        void* dm = (c.bf && dB[s].mlp.d_out) ? act(c, n, C) : nullptr;   // mask(D) for the block's MLP, written by the CA block's last LayerNorm backward
        ca_bwd(c, m.decCA[s], sg[s], D, ctx, dctx, &dctx_init, dC[s], dm, dB[s].mlp.d_out);
        void* dconv = nullptr;
        block_bwd(c, m.decB[s], sg[s], D, dB[s], &dconv, dm);        // D = d(xu); dconv = D + conv-input grad = d(b)
        // unpool: xu = b + a[cluster]
        void* da = act(c, nc, C);
        if (c.bf) CALL(lotus_segment_sum_bf16, D, C, sg[s + 1].seg_ptr, sg[s + 1].member, nc, 0, da, c.stv);
        else CALL(lotus_segment_sum_f32, (const float*)D, C, sg[s + 1].seg_ptr, sg[s + 1].member, nc, 0, (float*)da, c.stv);
        // skip branch: b = gelu(bn(lin(skip)))
        void* dlin_b = act(c, n, C);
        bn_bwd(c, dconv, dlin_b, m.unpool[s].bn_skip, upBnSkip[s]);
        Gskip[s] = act(c, n, m.encC[s]);
        linear_bwd(c, dlin_b, skip[s], n, m.unpool[s].skip, Gskip[s], nullptr, c.bf);
        // child branch: a = gelu(bn(lin(x_child)))
        void* dlin_a = act(c, nc, C);
        bn_bwd(c, da, dlin_a, m.unpool[s].bn, upBn[s]);
        int cin = m.unpool[s].proj.in;
        void* Dn = act(c, nc, cin);
        linear_bwd(c, dlin_a, up_in[s], nc, m.unpool[s].proj, Dn, nullptr, c.bf);
        D = Dn;                                                       // gradient of the deeper stage's residual stream
      }
      auto slice_done = [&](int k) {                                 // gradients of slice k are final on the main stream
        if (slice_events == nullptr || slice_events[k] == nullptr || !live(c)) return;
        join_side(c);
        cudaEventRecord((cudaEvent_t)slice_events[k], c.st);
      };
      slice_done(0);                                                 // head + decoder
      // encoder, deep to shallow
      for (int s = S - 1; s >= 0; --s) {
        int n = sg[s].n, C = m.encC[s];
        void* dm = (c.bf && eB[s].mlp.d_out) ? act(c, n, C) : nullptr;
        ca_bwd(c, m.encCA[s], sg[s], D, ctx, dctx, &dctx_init, eC[s], dm, eB[s].mlp.d_out);
        block_bwd(c, m.encB[s], sg[s], D, eB[s], nullptr, dm);        // D = d(stage input)
        if (s > 0) {
          int np = sg[s - 1].n;
          void* dpooled = act(c, n, C);
          bn_bwd(c, D, dpooled, m.pool[s].bn, poolBn[s]);
          void* dproj = act(c, np, C);
          zero(c, dproj, (size_t)np * C * c.es);
          if (c.bf) CALL(lotus_segment_max_bwd_bf16, dpooled, pool_arg[s], n, C, dproj, 0, c.stv);
          else CALL(lotus_segment_max_bwd_f32, (const float*)dpooled, pool_arg[s], n, C, (float*)dproj, 0, c.stv);
          // d(skip[s-1]) = Gskip[s-1] (decoder) + dproj @ Wpool
          linear_bwd(c, dproj, skip[s - 1], np, m.pool[s].proj, Gskip[s - 1], Gskip[s - 1], c.bf);
          D = Gskip[s - 1];
        }
        if (s == S - 1) slice_done(1);
        else if (s == S - 2) slice_done(2);
      }
      // stem: BN+GELU backward then the narrow conv weight gradient (the input features carry no gradient)
      {
        float* dstem = f32(c, (int64_t)N0 * m.encC[0]);
        bn_bwd(c, D, dstem, m.stem_bn, stem_sv);
        if (Gf(c, m.stem_w)) CALL(lotus_subm_conv_wgrad_f32, dstem, feat, Gf(c, m.stem_w), sg[0].nbr5, sg[0].stride, N0, m.inC, m.encC[0], 125, c.stv);
      }
      // context tokens
      {
        const void* txt_b = txt_in;
        linear_bwd(c, dctx, txt_b, Lt, m.txt_fc, nullptr, nullptr, c.bf);
      }
    }
  }
  join_side(c);                                    // the caller's stream owns every result from here on
  if (arena_peak) *arena_peak = (int64_t)c.ar.peak;
  if (!c.dry && (c.ar.overflow || c.ar.peak > c.ar.cap)) {
    lotus_set_error("lotus_policy_step: arena too small (%lld needed, %lld given)", (long long)c.ar.peak, (long long)c.ar.cap);
    return 3;
  }
  return c.rc;
}

// number of kernel-launching calls issued by the last (non-dry) lotus_policy_step on this process
LOTUS_API int64_t lotus_policy_step_calls(void) { return (int64_t)eng::g_launcher_calls; }

// Profile mode of lotus_policy_step: when enabled, the next step brackets every launcher call with CUDA events on the caller's
// stream (the weight-gradient side stream is folded back so the times are per call) and keeps (name, arguments).
LOTUS_API int lotus_policy_step_profile(int enable) { eng::g_prof_on = enable != 0; return 0; }
LOTUS_API int64_t lotus_policy_step_profile_count(void) { return (int64_t)eng::g_prof.size(); }
// record `index` of the last profiled step (synchronise the stream first): name (<= 63 chars + NUL), device milliseconds,
// the call's arguments in C-ABI order (pointers as addresses, floats as 0/1 = zero / non-zero), their count.  Host pointers.
LOTUS_API int lotus_policy_step_profile_read(int64_t index, char* name_out, float* ms_out, int64_t* args_out, int* nargs_out) {
  if (index < 0 || index >= (int64_t)eng::g_prof.size()) { lotus_set_error("lotus_policy_step_profile_read: index out of range"); return 2; }
  const eng::ProfRec& r = eng::g_prof[(size_t)index];
  strncpy(name_out, r.name, 63); name_out[63] = 0;
  float ms = 0.f;
  LOTUS_CUDA_CALL(cudaEventElapsedTime(&ms, r.e0, r.e1));
  *ms_out = ms; *nargs_out = r.nargs;
  for (int i = 0; i < r.nargs && i < 48; ++i) args_out[i] = r.a[i];
  return 0;
}
