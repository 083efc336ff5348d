// Attention kernels, fp32 math with fp32 or bf16 activations (flash-branch semantics of the reference):
//   patch self-attention over the serialized patches  (SerializedAttention.forward, model.py:468-557)
//   cross-attention of each cloud's points to its text tokens (CrossAttention.forward, model_ca.py:46-101)
// The gather by `order[pad]` and the scatter by `unpad[inverse]` are fused into the loads / stores, the
// optional relative-position bias (RPE, model.py:307-326) is added to the logits in-kernel, softmax is
// online, attention dropout uses the counter-based mask (recomputed in the backward).
// q and k arrive already per-head LayerNorm'ed (norm_act.cu).
#include "common.cuh"

// tensor-core variants (attention_tc.cu)
// warp-mma cross attention (cross_attn_mma.cu)
bool lotus_cross_attn_mma_ok(const void* q, const void* k, const void* v, const void* out, int64_t ldq, int64_t ldk, int64_t ldv,
                             int64_t ldo, int Dh, int max_tokens);
int lotus_cross_attn_fwd_mma(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldk, int64_t ldv, const int32_t* pt_off,
                             const int32_t* tok_off, int B, int max_points, int max_tokens, int H, int Dh, float scale, void* out,
                             int64_t ldo, float* lse, float drop_p, uint64_t seed, uint32_t layer_id, void* stream_);
int lotus_cross_attn_bwd_mma(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldk, int64_t ldv, const int32_t* pt_off,
                             const int32_t* tok_off, int B, int max_points, int max_tokens, int H, int Dh, float scale, const void* out,
                             const void* dout, int64_t ldo, const float* lse, float* dq, int64_t lddq, float* dk, float* dv,
                             int64_t lddk, int64_t lddv, float drop_p, uint64_t seed, uint32_t layer_id, void* stream_);


struct RpeArgs {
  const float* table;   // [3*(2*bnd+1), H] or null
  const int32_t* grid;  // [N,3]
  int pos_bnd;
  float* dtable;        // backward only
};

__device__ __forceinline__ float rpe_bias_ij(const RpeArgs& r, int H, int h, int gi, int gj) {
  int num = 2 * r.pos_bnd + 1;
  float b = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int d = r.grid[gi * 3 + a] - r.grid[gj * 3 + a];
    d = max(-r.pos_bnd, min(r.pos_bnd, d));
    b += r.table[(d + r.pos_bnd + a * num) * H + h];
  }
  return b;
}

// ---------------------------------------------------------------------------------------------- patch fwd
template <int DH, typename T>
__global__ void __launch_bounds__(128) k_patch_attn_fwd(
    const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, int64_t ldq, int64_t ldk,
    int64_t ldv, const int32_t* __restrict__ gather_idx, const int32_t* __restrict__ slot_of_point,
    const int32_t* __restrict__ cu, int H, float scale, RpeArgs rpe, T* __restrict__ out, int64_t ldo,
    float* __restrict__ lse, float drop_p, uint64_t seed, uint32_t layer_id) {
  extern __shared__ float sm[];
  int p = blockIdx.x, h = blockIdx.y;
  int s0 = cu[p], len = cu[p + 1] - s0;
  if (len <= 0) return;
  float* Ks = sm;
  float* Vs = Ks + (size_t)len * DH;
  int* Gs = (int*)(Vs + (size_t)len * DH);
  for (int t = threadIdx.x; t < len * DH; t += blockDim.x) {
    int j = t / DH, d = t - j * DH;
    int pt = gather_idx[s0 + j];
    Ks[t] = ldf(k, (int64_t)pt * ldk + h * DH + d);
    Vs[t] = ldf(v, (int64_t)pt * ldv + h * DH + d);
  }
  for (int j = threadIdx.x; j < len; j += blockDim.x) Gs[j] = gather_idx[s0 + j];
  __syncthreads();
  float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    int pt = Gs[i];
    float qr[DH], o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) { qr[d] = ldf(q, (int64_t)pt * ldq + h * DH + d) * scale; o[d] = 0.f; }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < len; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) s = fmaf(qr[d], Ks[j * DH + d], s);
      if (rpe.table) s += rpe_bias_ij(rpe, H, h, pt, Gs[j]);
      float mn = fmaxf(m, s);
      float corr = __expf(m - mn), pj = __expf(s - mn);
      l = l * corr + pj;
      float pv = pj;
      if (drop_p > 0.f)
        pv = attn_drop_keep(attn_drop_rowhash(seed, layer_id * 0x9E37u + (uint32_t)(p * H + h), (uint32_t)i), (uint32_t)j, attn_drop_thr(drop_p)) ? pj * keep_scale : 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = o[d] * corr + pv * Vs[j * DH + d];
      m = mn;
    }
    float inv = 1.f / l;
    lse[(int64_t)(s0 + i) * H + h] = m + __logf(l);
    if (slot_of_point[pt] == s0 + i) {  // alias slots of a padded tail are queries whose output is discarded
#pragma unroll
      for (int d = 0; d < DH; ++d) stf(out, (int64_t)pt * ldo + h * DH + d, o[d] * inv);
    }
  }
}

// ---------------------------------------------------------------------------------------------- patch bwd
template <int DH, typename T>
__global__ void __launch_bounds__(128) k_patch_attn_bwd(
    const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, int64_t ldq, int64_t ldk,
    int64_t ldv, const int32_t* __restrict__ gather_idx, const int32_t* __restrict__ slot_of_point,
    const int32_t* __restrict__ cu, int H, float scale, RpeArgs rpe, const T* __restrict__ out,
    const T* __restrict__ dout, int64_t ldo, const float* __restrict__ lse, float* __restrict__ dq,
    float* __restrict__ dk, float* __restrict__ dv, int64_t lddq, int64_t lddk, int64_t lddv, float drop_p,
    uint64_t seed, uint32_t layer_id) {
  extern __shared__ float sm[];
  int p = blockIdx.x, h = blockIdx.y;
  int s0 = cu[p], len = cu[p + 1] - s0;
  if (len <= 0) return;
  float* Qs = sm;
  float* Ks = Qs + (size_t)len * DH;
  float* Vs = Ks + (size_t)len * DH;
  float* dOs = Vs + (size_t)len * DH;
  float* Ls = dOs + (size_t)len * DH;
  float* Ds = Ls + len;
  int* Gs = (int*)(Ds + len);
  for (int j = threadIdx.x; j < len; j += blockDim.x) Gs[j] = gather_idx[s0 + j];
  __syncthreads();
  for (int t = threadIdx.x; t < len * DH; t += blockDim.x) {
    int j = t / DH, d = t - j * DH;
    int pt = Gs[j];
    bool real = slot_of_point[pt] == s0 + j;
    Qs[t] = ldf(q, (int64_t)pt * ldq + h * DH + d) * scale;
    Ks[t] = ldf(k, (int64_t)pt * ldk + h * DH + d);
    Vs[t] = ldf(v, (int64_t)pt * ldv + h * DH + d);
    dOs[t] = real ? ldf(dout, (int64_t)pt * ldo + h * DH + d) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    int pt = Gs[i];
    float dsum = 0.f;
    if (slot_of_point[pt] == s0 + i) {
#pragma unroll
      for (int d = 0; d < DH; ++d) dsum = fmaf(dOs[i * DH + d], ldf(out, (int64_t)pt * ldo + h * DH + d), dsum);
    }
    Ds[i] = dsum;
    Ls[i] = lse[(int64_t)(s0 + i) * H + h];
  }
  __syncthreads();
  float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  uint32_t sid = layer_id * 0x9E37u + (uint32_t)(p * H + h);
  // phase B: thread = key j -> dK_j, dV_j
  for (int j = threadIdx.x; j < len; j += blockDim.x) {
    float kr[DH], vr[DH], ak[DH], av[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) { kr[d] = Ks[j * DH + d]; vr[d] = Vs[j * DH + d]; ak[d] = 0.f; av[d] = 0.f; }
    for (int i = 0; i < len; ++i) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) { s = fmaf(Qs[i * DH + d], kr[d], s); dp = fmaf(dOs[i * DH + d], vr[d], dp); }
      if (rpe.table) s += rpe_bias_ij(rpe, H, h, Gs[i], Gs[j]);
      float pij = __expf(s - Ls[i]);
      float pd = pij;
      if (drop_p > 0.f) {
        bool keep = attn_drop_keep(attn_drop_rowhash(seed, sid, (uint32_t)i), (uint32_t)j, attn_drop_thr(drop_p));
        pd = keep ? pij * keep_scale : 0.f;
        dp = keep ? dp * keep_scale : 0.f;
      }
      float ds = pij * (dp - Ds[i]);
#pragma unroll
      for (int d = 0; d < DH; ++d) { av[d] = fmaf(pd, dOs[i * DH + d], av[d]); ak[d] = fmaf(ds, Qs[i * DH + d], ak[d]); }
    }
    int pt = Gs[j];
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      atomicAdd(&dk[(int64_t)pt * lddk + h * DH + d], ak[d]);   // Qs already carries `scale`
      atomicAdd(&dv[(int64_t)pt * lddv + h * DH + d], av[d]);
    }
  }
  // phase C: thread = query i -> dQ_i (and the RPE table gradient)
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    float qr[DH], dor[DH], aq[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) { qr[d] = Qs[i * DH + d]; dor[d] = dOs[i * DH + d]; aq[d] = 0.f; }
    float li = Ls[i], di = Ds[i];
    for (int j = 0; j < len; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) { s = fmaf(qr[d], Ks[j * DH + d], s); dp = fmaf(dor[d], Vs[j * DH + d], dp); }
      if (rpe.table) s += rpe_bias_ij(rpe, H, h, Gs[i], Gs[j]);
      float pij = __expf(s - li);
      if (drop_p > 0.f) dp = attn_drop_keep(attn_drop_rowhash(seed, sid, (uint32_t)i), (uint32_t)j, attn_drop_thr(drop_p)) ? dp * keep_scale : 0.f;
      float ds = pij * (dp - di);
#pragma unroll
      for (int d = 0; d < DH; ++d) aq[d] = fmaf(ds, Ks[j * DH + d], aq[d]);
      if (rpe.dtable && ds != 0.f) {
        int num = 2 * rpe.pos_bnd + 1;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          int dd = rpe.grid[Gs[i] * 3 + a] - rpe.grid[Gs[j] * 3 + a];
          dd = max(-rpe.pos_bnd, min(rpe.pos_bnd, dd));
          atomicAdd(&rpe.dtable[(dd + rpe.pos_bnd + a * num) * H + h], ds);
        }
      }
    }
    int pt = Gs[i];
#pragma unroll
    for (int d = 0; d < DH; ++d) atomicAdd(&dq[(int64_t)pt * lddq + h * DH + d], aq[d] * scale);
  }
}

#define DISPATCH_DH(DHV, ...)                                   \
  switch (DHV) {                                                \
    case 8: { constexpr int DH = 8; __VA_ARGS__; } break;       \
    case 16: { constexpr int DH = 16; __VA_ARGS__; } break;     \
    case 24: { constexpr int DH = 24; __VA_ARGS__; } break;     \
    case 32: { constexpr int DH = 32; __VA_ARGS__; } break;     \
    case 48: { constexpr int DH = 48; __VA_ARGS__; } break;     \
    case 64: { constexpr int DH = 64; __VA_ARGS__; } break;     \
    default:                                                    \
      lotus_set_error("attention: head_dim %d unsupported (8,16,24,32,48,64)", DHV); \
      return 2;                                                 \
  }

template <typename T>
static int patch_fwd_impl(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                          const int32_t* gather_idx, const int32_t* slot_of_point, const int32_t* cu_seqlens, int n_patches,
                          int max_len, int H, int Dh, float scale, const float* rpe_table, const int32_t* grid, int pos_bnd,
                          void* out, int64_t ldo, float* lse, float drop_p, uint64_t seed, uint32_t layer_id, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_patches <= 0) return 0;
  size_t smem = (size_t)max_len * Dh * 8 + (size_t)max_len * 4;
  LOTUS_CHECK_ARG(smem <= 200 * 1024, "lotus_patch_attn_fwd: patch %d x head_dim %d exceeds shared memory", max_len, Dh);
  RpeArgs rpe{rpe_table, grid, pos_bnd, nullptr};
  DISPATCH_DH(Dh, {
    LOTUS_CUDA_CALL(cudaFuncSetAttribute(k_patch_attn_fwd<DH, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_patch_attn_fwd<DH, T><<<dim3(n_patches, H), 128, smem, st>>>((const T*)q, (const T*)k, (const T*)v, ldq, ldk, ldv, gather_idx,
                                                                   slot_of_point, cu_seqlens, H, scale, rpe, (T*)out, ldo, lse,
                                                                   drop_p, seed, layer_id);
  });
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_patch_attn_fwd_f32(const float* q, const float* k, const float* v, int64_t ldq, int64_t ldk,
                                       int64_t ldv, const int32_t* gather_idx, const int32_t* slot_of_point,
                                       const int32_t* cu_seqlens, int n_patches, int max_len, int H, int Dh, float scale,
                                       const float* rpe_table, const int32_t* grid, int pos_bnd, float* out,
                                       int64_t ldo, float* lse, float drop_p, uint64_t seed, uint32_t layer_id,
                                       void* stream_) {
  return patch_fwd_impl<float>(q, k, v, ldq, ldk, ldv, gather_idx, slot_of_point, cu_seqlens, n_patches, max_len, H, Dh, scale,
                               rpe_table, grid, pos_bnd, out, ldo, lse, drop_p, seed, layer_id, stream_);
}
// q, k, v, out bf16 (fp32 softmax / accumulation inside)
LOTUS_API int lotus_patch_attn_fwd_bf16(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                                        const int32_t* gather_idx, const int32_t* slot_of_point, const int32_t* cu_seqlens,
                                        int n_patches, int max_len, int H, int Dh, float scale, const float* rpe_table,
                                        const int32_t* grid, int pos_bnd, void* out, int64_t ldo, float* lse, float drop_p,
                                        uint64_t seed, uint32_t layer_id, void* stream_) {
  // (bf16 activations through the SIMT kernel: A/B reference of the fused tcgen05 kernels in attention_fused.cu, and shapes those refuse)
  return patch_fwd_impl<bf16_t>(q, k, v, ldq, ldk, ldv, gather_idx, slot_of_point, cu_seqlens, n_patches, max_len, H, Dh, scale,
                                rpe_table, grid, pos_bnd, out, ldo, lse, drop_p, seed, layer_id, stream_);
}

// dq/dk/dv (fp32) are ACCUMULATED with atomics (alias slots add a second contribution): zero them first.
template <typename T>
static int patch_bwd_impl(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                          const int32_t* gather_idx, const int32_t* slot_of_point, const int32_t* cu_seqlens, int n_patches,
                          int max_len, int H, int Dh, float scale, const float* rpe_table, const int32_t* grid, int pos_bnd,
                          float* drpe_table, const void* out, const void* dout, int64_t ldo, const float* lse, float* dq,
                          float* dk, float* dv, int64_t lddq, int64_t lddk, int64_t lddv, float drop_p, uint64_t seed,
                          uint32_t layer_id, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_patches <= 0) return 0;
  size_t smem = (size_t)max_len * Dh * 16 + (size_t)max_len * 12;
  LOTUS_CHECK_ARG(smem <= 200 * 1024, "lotus_patch_attn_bwd: patch %d x head_dim %d exceeds shared memory", max_len, Dh);
  RpeArgs rpe{rpe_table, grid, pos_bnd, drpe_table};
  DISPATCH_DH(Dh, {
    LOTUS_CUDA_CALL(cudaFuncSetAttribute(k_patch_attn_bwd<DH, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_patch_attn_bwd<DH, T><<<dim3(n_patches, H), 128, smem, st>>>((const T*)q, (const T*)k, (const T*)v, ldq, ldk, ldv, gather_idx,
                                                                   slot_of_point, cu_seqlens, H, scale, rpe, (const T*)out,
                                                                   (const T*)dout, ldo, lse, dq, dk, dv, lddq, lddk, lddv, drop_p,
                                                                   seed, layer_id);
  });
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_patch_attn_bwd_f32(const float* q, const float* k, const float* v, int64_t ldq, int64_t ldk,
                                       int64_t ldv, const int32_t* gather_idx, const int32_t* slot_of_point,
                                       const int32_t* cu_seqlens, int n_patches, int max_len, int H, int Dh, float scale,
                                       const float* rpe_table, const int32_t* grid, int pos_bnd, float* drpe_table,
                                       const float* out, const float* dout, int64_t ldo, const float* lse, float* dq,
                                       float* dk, float* dv, int64_t lddq, int64_t lddk, int64_t lddv, float drop_p,
                                       uint64_t seed, uint32_t layer_id, void* stream_) {
  return patch_bwd_impl<float>(q, k, v, ldq, ldk, ldv, gather_idx, slot_of_point, cu_seqlens, n_patches, max_len, H, Dh, scale,
                               rpe_table, grid, pos_bnd, drpe_table, out, dout, ldo, lse, dq, dk, dv, lddq, lddk, lddv, drop_p, seed,
                               layer_id, stream_);
}
// q, k, v, out, dout bf16; dq, dk, dv fp32 (accumulated)
LOTUS_API int lotus_patch_attn_bwd_bf16(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                                        const int32_t* gather_idx, const int32_t* slot_of_point, const int32_t* cu_seqlens,
                                        int n_patches, int max_len, int H, int Dh, float scale, const float* rpe_table,
                                        const int32_t* grid, int pos_bnd, float* drpe_table, const void* out, const void* dout,
                                        int64_t ldo, const float* lse, float* dq, float* dk, float* dv, int64_t lddq,
                                        int64_t lddk, int64_t lddv, float drop_p, uint64_t seed, uint32_t layer_id,
                                        void* stream_) {
  return patch_bwd_impl<bf16_t>(q, k, v, ldq, ldk, ldv, gather_idx, slot_of_point, cu_seqlens, n_patches, max_len, H, Dh, scale,
                                rpe_table, grid, pos_bnd, drpe_table, out, dout, ldo, lse, dq, dk, dv, lddq, lddk, lddv, drop_p,
                                seed, layer_id, stream_);
}

// ---------------------------------------------------------------------------------------------- cross attention
template <int DH, typename T>
__global__ void __launch_bounds__(128) k_cross_attn_fwd(
    const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, const T* __restrict__ v, int64_t ldk,
    int64_t ldv, const int32_t* __restrict__ pt_off, const int32_t* __restrict__ tok_off, int H, float scale,
    T* __restrict__ out, int64_t ldo, float* __restrict__ lse, float drop_p, uint64_t seed, uint32_t layer_id) {
  extern __shared__ float sm[];
  int b = blockIdx.y, h = blockIdx.z;
  int ps = pt_off[b], pn = pt_off[b + 1] - ps;
  if ((int)(blockIdx.x * blockDim.x) >= pn) return;
  int ts = tok_off[b], L = tok_off[b + 1] - ts;
  float* Ks = sm;
  float* Vs = Ks + (size_t)L * DH;
  for (int t = threadIdx.x; t < L * DH; t += blockDim.x) {
    int j = t / DH, d = t - j * DH;
    Ks[t] = ldf(k, (int64_t)(ts + j) * ldk + h * DH + d);
    Vs[t] = ldf(v, (int64_t)(ts + j) * ldv + h * DH + d);
  }
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pn) return;
  int pt = ps + i;
  float qr[DH], o[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) { qr[d] = ldf(q, (int64_t)pt * ldq + h * DH + d) * scale; o[d] = 0.f; }
  float m = -INFINITY, l = 0.f;
  float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (int j = 0; j < L; ++j) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) s = fmaf(qr[d], Ks[j * DH + d], s);
    float mn = fmaxf(m, s);
    float corr = __expf(m - mn), pj = __expf(s - mn);
    l = l * corr + pj;
    float pv = pj;
    if (drop_p > 0.f) pv = dropout_keep(seed, layer_id * 0x9E37u + (uint32_t)h, (uint32_t)(pt * 128 + j), drop_p) ? pj * keep_scale : 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = o[d] * corr + pv * Vs[j * DH + d];
    m = mn;
  }
  float inv = L > 0 ? 1.f / l : 0.f;
  lse[(int64_t)pt * H + h] = m + __logf(l);
#pragma unroll
  for (int d = 0; d < DH; ++d) stf(out, (int64_t)pt * ldo + h * DH + d, o[d] * inv);
}

template <int DH, typename T>
__global__ void __launch_bounds__(128) k_cross_attn_bwd(
    const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, const T* __restrict__ v, int64_t ldk,
    int64_t ldv, const int32_t* __restrict__ pt_off, const int32_t* __restrict__ tok_off, int H, float scale,
    const T* __restrict__ out, const T* __restrict__ dout, int64_t ldo, const float* __restrict__ lse,
    float* __restrict__ dq, int64_t lddq, float* __restrict__ dk, float* __restrict__ dv, int64_t lddk, int64_t lddv,
    float drop_p, uint64_t seed, uint32_t layer_id) {
  extern __shared__ float sm[];
  int b = blockIdx.y, h = blockIdx.z;
  int ps = pt_off[b], pn = pt_off[b + 1] - ps;
  int i0 = blockIdx.x * blockDim.x;
  if (i0 >= pn) return;
  int cn = min((int)blockDim.x, pn - i0);  // points in this chunk
  int ts = tok_off[b], L = tok_off[b + 1] - ts;
  float* Ks = sm;
  float* Vs = Ks + (size_t)L * DH;
  float* Qs = Vs + (size_t)L * DH;          // [128][DH]  (scaled)
  float* dOs = Qs + (size_t)128 * DH;       // [128][DH]
  float* Ls = dOs + (size_t)128 * DH;       // [128]
  float* Ds = Ls + 128;                     // [128]
  const int LP = L + 1;                     // padded row stride of the P / dS tiles
  float* Ps = Ds + 128;                     // [128][LP] dropped probabilities (for dV)
  float* dSs = Ps + (size_t)128 * LP;       // [128][LP] score gradients (for dK)
  for (int t = threadIdx.x; t < L * DH; t += blockDim.x) {
    int j = t / DH, d = t - j * DH;
    Ks[t] = ldf(k, (int64_t)(ts + j) * ldk + h * DH + d);
    Vs[t] = ldf(v, (int64_t)(ts + j) * ldv + h * DH + d);
  }
  for (int t = threadIdx.x; t < cn * DH; t += blockDim.x) {
    int i = t / DH, d = t - i * DH;
    int pt = ps + i0 + i;
    Qs[t] = ldf(q, (int64_t)pt * ldq + h * DH + d) * scale;
    dOs[t] = ldf(dout, (int64_t)pt * ldo + h * DH + d);
  }
  __syncthreads();
  float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  uint32_t sid = layer_id * 0x9E37u + (uint32_t)h;
  int i = threadIdx.x;
  if (i < cn) {
    int pt = ps + i0 + i;
    float di = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) di = fmaf(dOs[i * DH + d], ldf(out, (int64_t)pt * ldo + h * DH + d), di);
    Ds[i] = di;
    Ls[i] = lse[(int64_t)pt * H + h];
    // dQ_i
    float qr[DH], dor[DH], aq[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) { qr[d] = Qs[i * DH + d]; dor[d] = dOs[i * DH + d]; aq[d] = 0.f; }
    float li = Ls[i];
    for (int j = 0; j < L; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) { s = fmaf(qr[d], Ks[j * DH + d], s); dp = fmaf(dor[d], Vs[j * DH + d], dp); }
      float pij = __expf(s - li), pd = pij;
      if (drop_p > 0.f) {
        bool keep = dropout_keep(seed, sid, (uint32_t)(pt * 128 + j), drop_p);
        pd = keep ? pij * keep_scale : 0.f;
        dp = keep ? dp * keep_scale : 0.f;
      }
      float ds = pij * (dp - di);
      Ps[i * LP + j] = pd;
      dSs[i * LP + j] = ds;
#pragma unroll
      for (int d = 0; d < DH; ++d) aq[d] = fmaf(ds, Ks[j * DH + d], aq[d]);
    }
#pragma unroll
    for (int d = 0; d < DH; ++d) dq[(int64_t)pt * lddq + h * DH + d] = aq[d] * scale;
  }
  __syncthreads();
  // dK, dV: thread = (token j, channel d) pairs over the chunk's points, from the P / dS tiles of phase 1
  for (int o = threadIdx.x; o < L * DH; o += blockDim.x) {
    int j = o / DH, d = o - j * DH;
    float ak = 0.f, av = 0.f;
    for (int ii = 0; ii < cn; ++ii) {
      ak = fmaf(dSs[ii * LP + j], Qs[ii * DH + d], ak);      // Qs already carries `scale`
      av = fmaf(Ps[ii * LP + j], dOs[ii * DH + d], av);
    }
    atomicAdd(&dk[(int64_t)(ts + j) * lddk + h * DH + d], ak);
    atomicAdd(&dv[(int64_t)(ts + j) * lddv + h * DH + d], av);
  }
}

template <typename T>
static int cross_fwd_impl(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldk, int64_t ldv,
                          const int32_t* pt_off, const int32_t* tok_off, int B, int max_points, int max_tokens, int H, int Dh,
                          float scale, void* out, int64_t ldo, float* lse, float drop_p, uint64_t seed, uint32_t layer_id,
                          void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (B <= 0 || max_points <= 0) return 0;
  size_t smem = (size_t)max_tokens * Dh * 8;
  LOTUS_CHECK_ARG(smem <= 200 * 1024, "lotus_cross_attn_fwd: %d tokens x head_dim %d exceeds shared memory", max_tokens, Dh);
  LOTUS_CHECK_ARG(max_tokens <= 128, "lotus_cross_attn_fwd: at most 128 context tokens per cloud (got %d)", max_tokens);
  dim3 grid(ceil_div(max_points, 128), B, H);
  DISPATCH_DH(Dh, {
    LOTUS_CUDA_CALL(cudaFuncSetAttribute(k_cross_attn_fwd<DH, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_cross_attn_fwd<DH, T><<<grid, 128, smem, st>>>((const T*)q, ldq, (const T*)k, (const T*)v, ldk, ldv, pt_off, tok_off, H, scale,
                                                     (T*)out, ldo, lse, drop_p, seed, layer_id);
  });
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_cross_attn_fwd_f32(const float* q, int64_t ldq, const float* k, const float* v, int64_t ldk,
                                       int64_t ldv, const int32_t* pt_off, const int32_t* tok_off, int B, int max_points,
                                       int max_tokens, int H, int Dh, float scale, float* out, int64_t ldo, float* lse,
                                       float drop_p, uint64_t seed, uint32_t layer_id, void* stream_) {
  return cross_fwd_impl<float>(q, ldq, k, v, ldk, ldv, pt_off, tok_off, B, max_points, max_tokens, H, Dh, scale, out, ldo, lse,
                               drop_p, seed, layer_id, stream_);
}
LOTUS_API int lotus_cross_attn_fwd_bf16(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldk, int64_t ldv,
                                        const int32_t* pt_off, const int32_t* tok_off, int B, int max_points, int max_tokens,
                                        int H, int Dh, float scale, void* out, int64_t ldo, float* lse, float drop_p,
                                        uint64_t seed, uint32_t layer_id, void* stream_) {
  if (lotus_cross_attn_mma_ok(q, k, v, out, ldq, ldk, ldv, ldo, Dh, max_tokens))                 // warp-mma path (cross_attn_mma.cu)
    return lotus_cross_attn_fwd_mma(q, ldq, k, v, ldk, ldv, pt_off, tok_off, B, max_points, max_tokens, H, Dh, scale, out, ldo, lse,
                                    drop_p, seed, layer_id, stream_);
  return cross_fwd_impl<bf16_t>(q, ldq, k, v, ldk, ldv, pt_off, tok_off, B, max_points, max_tokens, H, Dh, scale, out, ldo, lse,
                                drop_p, seed, layer_id, stream_);
}

// dq (fp32) is written; dk/dv (fp32) are ACCUMULATED with atomics: zero them first.
template <typename T>
static int cross_bwd_impl(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldk, int64_t ldv,
                          const int32_t* pt_off, const int32_t* tok_off, int B, int max_points, int max_tokens, int H, int Dh,
                          float scale, const void* out, const void* dout, int64_t ldo, const float* lse, float* dq,
                          int64_t lddq, float* dk, float* dv, int64_t lddk, int64_t lddv, float drop_p, uint64_t seed,
                          uint32_t layer_id, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (B <= 0 || max_points <= 0) return 0;
  size_t smem = (size_t)max_tokens * Dh * 8 + (size_t)128 * Dh * 8 + 1024 + (size_t)128 * (max_tokens + 1) * 8;
  LOTUS_CHECK_ARG(smem <= 200 * 1024, "lotus_cross_attn_bwd: %d tokens x head_dim %d exceeds shared memory", max_tokens, Dh);
  dim3 grid(ceil_div(max_points, 128), B, H);
  DISPATCH_DH(Dh, {
    LOTUS_CUDA_CALL(cudaFuncSetAttribute(k_cross_attn_bwd<DH, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_cross_attn_bwd<DH, T><<<grid, 128, smem, st>>>((const T*)q, ldq, (const T*)k, (const T*)v, ldk, ldv, pt_off, tok_off, H, scale,
                                                     (const T*)out, (const T*)dout, ldo, lse, dq, lddq, dk, dv, lddk, lddv, drop_p,
                                                     seed, layer_id);
  });
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_cross_attn_bwd_f32(const float* q, int64_t ldq, const float* k, const float* v, int64_t ldk,
                                       int64_t ldv, const int32_t* pt_off, const int32_t* tok_off, int B, int max_points,
                                       int max_tokens, int H, int Dh, float scale, const float* out, const float* dout,
                                       int64_t ldo, const float* lse, float* dq, int64_t lddq, float* dk, float* dv,
                                       int64_t lddk, int64_t lddv, float drop_p, uint64_t seed, uint32_t layer_id,
                                       void* stream_) {
  return cross_bwd_impl<float>(q, ldq, k, v, ldk, ldv, pt_off, tok_off, B, max_points, max_tokens, H, Dh, scale, out, dout, ldo, lse,
                               dq, lddq, dk, dv, lddk, lddv, drop_p, seed, layer_id, stream_);
}
LOTUS_API int lotus_cross_attn_bwd_bf16(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldk, int64_t ldv,
                                        const int32_t* pt_off, const int32_t* tok_off, int B, int max_points, int max_tokens,
                                        int H, int Dh, float scale, const void* out, const void* dout, int64_t ldo,
                                        const float* lse, float* dq, int64_t lddq, float* dk, float* dv, int64_t lddk,
                                        int64_t lddv, float drop_p, uint64_t seed, uint32_t layer_id, void* stream_) {
  if (lotus_cross_attn_mma_ok(q, k, v, out, ldq, ldk, ldv, ldo, Dh, max_tokens) && ((uintptr_t)dout & 3) == 0 &&
      ((uintptr_t)dq & 7) == 0 && lddq % 2 == 0)
    return lotus_cross_attn_bwd_mma(q, ldq, k, v, ldk, ldv, pt_off, tok_off, B, max_points, max_tokens, H, Dh, scale, out, dout, ldo,
                                    lse, dq, lddq, dk, dv, lddk, lddv, drop_p, seed, layer_id, stream_);
  return cross_bwd_impl<bf16_t>(q, ldq, k, v, ldk, ldv, pt_off, tok_off, B, max_points, max_tokens, H, Dh, scale, out, dout, ldo,
                                lse, dq, lddq, dk, dv, lddk, lddv, drop_p, seed, layer_id, stream_);
}
