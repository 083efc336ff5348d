// Grid pooling / unpooling reductions, the action head's per-cloud reductions, the losses and the
// final-action decode, fp32.
//   segment max / mean over CSR clusters ........ torch_scatter.segment_csr, model.py:760-765
//   unpool gather-add ........................... model.py:824
//   per-cloud max over points ................... simple_policy_ptv3.py:117-119
//   soft-label CE + rot CE + open BCE ........... simple_policy_ptv3.py:308-373
//   argmax decode + euler bins -> quaternion .... simple_policy_ptv3.py:250-295, action_position_utils.py:48-60,
//                                                 rotation_transform.py:192-193
#include "common.cuh"
#include <cooperative_groups.h>

// out[m, c] = max_{j in [ptr[m], ptr[m+1])} x[member ? member[j] : j, c]; arg[m, c] = winning source row
template <typename T>
__global__ void k_segment_max(const T* __restrict__ x, int C, const int32_t* __restrict__ ptr,
                              const int32_t* __restrict__ member, int M, T* __restrict__ out,
                              int32_t* __restrict__ arg) {
  int64_t total = (int64_t)M * C;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int m = (int)(t / C), c = (int)(t % C);
    int s = ptr[m], e = ptr[m + 1];
    float best = -INFINITY; int bi = -1;
    for (int j = s; j < e; ++j) {
      int r = member ? member[j] : j;
      float v = ldf(x, (int64_t)r * C + c);
      if (v > best || bi < 0) { best = v; bi = r; }
    }
    stf(out, t, best);
    if (arg) arg[t] = bi;
  }
}
// Long contiguous segments (the per-cloud max of the action head: M = clouds, thousands of rows each):
// one CTA per (segment, 8-channel chunk); 256 threads stride over the rows with 16/32-byte loads, then a
// shared-memory tree reduction.  Ties -> lowest row, like the sequential kernel above.
template <typename T>
__global__ void __launch_bounds__(256) k_segment_max_long(const T* __restrict__ x, int C, const int32_t* __restrict__ ptr, T* __restrict__ out,
                                                          int32_t* __restrict__ arg) {
  __shared__ float sv[256][8];
  __shared__ int si[256][8];
  const int m = blockIdx.x, c0 = blockIdx.y * 8, tid = threadIdx.x;
  const int s = ptr[m], e = ptr[m + 1];
  float best[8]; int bi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { best[i] = -INFINITY; bi[i] = -1; }
  for (int r = s + tid; r < e; r += 256) {
    float v[8];
    if constexpr (sizeof(T) == 2) {
      uint4 raw = *reinterpret_cast<const uint4*>(x + (int64_t)r * C + c0);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    } else {
      const float4* p4 = reinterpret_cast<const float4*>(x + (int64_t)r * C + c0);
      float4 a = p4[0], b = p4[1];
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) if (v[i] > best[i] || bi[i] < 0) { best[i] = v[i]; bi[i] = r; }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { sv[tid][i] = best[i]; si[tid][i] = bi[i]; }
  __syncthreads();
  for (int half = 128; half >= 1; half >>= 1) {
    if (tid < half) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float vo = sv[tid + half][i]; int io = si[tid + half][i];
        float vm = sv[tid][i]; int im = si[tid][i];
        // the other half wins on a larger value, or on an equal value from a lower row; empty slots (-1) never win
        bool take = io >= 0 && (im < 0 || vo > vm || (vo == vm && io < im));
        if (take) { sv[tid][i] = vo; si[tid][i] = io; }
      }
    }
    __syncthreads();
  }
  if (tid < 8) {
    int64_t o = (int64_t)m * C + c0 + tid;
    stf(out, o, sv[0][tid]);
    if (arg) arg[o] = si[0][tid];
  }
}
template <typename T>
static int segmax_fwd_impl(const void* x, int C, const int32_t* seg_ptr, const int32_t* member, int M, void* out, int32_t* argmax,
                           void* stream_) {
  if (M <= 0) return 0;
  int64_t total = (int64_t)M * C;
  if (member == nullptr && M <= 1024 && C % 8 == 0) {
    k_segment_max_long<T><<<dim3(M, C / 8), 256, 0, (cudaStream_t)stream_>>>((const T*)x, C, seg_ptr, (T*)out, argmax);
    LOTUS_CUDA_OK();
    return 0;
  }
  k_segment_max<T><<<min(ceil_div(total, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>((const T*)x, C, seg_ptr, member, M,
                                                                                             (T*)out, argmax);
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_segment_max_fwd_f32(const float* x, int C, const int32_t* seg_ptr, const int32_t* member, int M,
                                        float* out, int32_t* argmax, void* stream_) {
  return segmax_fwd_impl<float>(x, C, seg_ptr, member, M, out, argmax, stream_);
}
LOTUS_API int lotus_segment_max_fwd_bf16(const void* x, int C, const int32_t* seg_ptr, const int32_t* member, int M, void* out,
                                         int32_t* argmax, void* stream_) {
  return segmax_fwd_impl<bf16_t>(x, C, seg_ptr, member, M, out, argmax, stream_);
}
// dx[arg[m,c], c] (+)= dy[m,c]   (segments are disjoint -> no atomics; accumulate=0 expects a zeroed dx)
template <typename T>
__global__ void k_segment_max_bwd(const T* __restrict__ dy, const int32_t* __restrict__ arg, int64_t total, int C,
                                  T* __restrict__ dx, int accumulate) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int r = arg[t];
    if (r >= 0) {
      int64_t o = (int64_t)r * C + (t % C);
      stf(dx, o, accumulate ? ldf(dx, o) + ldf(dy, t) : ldf(dy, t));
    }
  }
}
LOTUS_API int lotus_segment_max_bwd_f32(const float* dy, const int32_t* argmax, int M, int C, float* dx, int accumulate,
                                        void* stream_) {
  if (M <= 0) return 0;
  int64_t total = (int64_t)M * C;
  k_segment_max_bwd<float><<<min(ceil_div(total, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>(dy, argmax, total, C, dx, accumulate);
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_segment_max_bwd_bf16(const void* dy, const int32_t* argmax, int M, int C, void* dx, int accumulate,
                                         void* stream_) {
  if (M <= 0) return 0;
  int64_t total = (int64_t)M * C;
  k_segment_max_bwd<bf16_t><<<min(ceil_div(total, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>((const bf16_t*)dy, argmax, total,
                                                                                                     C, (bf16_t*)dx, accumulate);
  LOTUS_CUDA_OK();
  return 0;
}

// op 0: mean, 1: sum  (sum is the backward of the unpool gather: d_child[m] = sum_{p in cluster m} dy[p])
template <typename T>
__global__ void k_segment_sum(const T* __restrict__ x, int C, const int32_t* __restrict__ ptr,
                              const int32_t* __restrict__ member, int M, int mean, T* __restrict__ out) {
  int64_t total = (int64_t)M * C;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int m = (int)(t / C), c = (int)(t % C);
    int s = ptr[m], e = ptr[m + 1];
    float acc = 0.f;
    for (int j = s; j < e; ++j) acc += ldf(x, (int64_t)(member ? member[j] : j) * C + c);
    stf(out, t, mean ? acc / (float)(e - s) : acc);
  }
}
LOTUS_API int lotus_segment_sum_f32(const float* x, int C, const int32_t* seg_ptr, const int32_t* member, int M,
                                    int mean, float* out, void* stream_) {
  if (M <= 0) return 0;
  int64_t total = (int64_t)M * C;
  k_segment_sum<float><<<min(ceil_div(total, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>(x, C, seg_ptr, member, M, mean, out);
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_segment_sum_bf16(const void* x, int C, const int32_t* seg_ptr, const int32_t* member, int M, int mean,
                                     void* out, void* stream_) {
  if (M <= 0) return 0;
  int64_t total = (int64_t)M * C;
  k_segment_sum<bf16_t><<<min(ceil_div(total, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>((const bf16_t*)x, C, seg_ptr, member,
                                                                                                  M, mean, (bf16_t*)out);
  LOTUS_CUDA_OK();
  return 0;
}

// out[p] = skip[p] + child[cluster[p]]
template <typename T>
__global__ void k_unpool_add(const T* __restrict__ skip, const T* __restrict__ child,
                             const int32_t* __restrict__ cluster, int64_t total, int C, T* __restrict__ out) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int p = (int)(t / C), c = (int)(t % C);
    stf(out, t, ldf(skip, t) + ldf(child, (int64_t)cluster[p] * C + c));
  }
}
LOTUS_API int lotus_unpool_add_f32(const float* skip, const float* child, const int32_t* cluster, int n, int C,
                                   float* out, void* stream_) {
  if (n <= 0) return 0;
  int64_t total = (int64_t)n * C;
  k_unpool_add<float><<<min(ceil_div(total, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>(skip, child, cluster, total, C, out);
  LOTUS_CUDA_OK();
  return 0;
}
LOTUS_API int lotus_unpool_add_bf16(const void* skip, const void* child, const int32_t* cluster, int n, int C, void* out,
                                    void* stream_) {
  if (n <= 0) return 0;
  int64_t total = (int64_t)n * C;
  k_unpool_add<bf16_t><<<min(ceil_div(total, 256), kNumSM * 16), 256, 0, (cudaStream_t)stream_>>>((const bf16_t*)skip, (const bf16_t*)child,
                                                                                                 cluster, total, C, (bf16_t*)out);
  LOTUS_CUDA_OK();
  return 0;
}

// ---------------------------------------------------------------- losses
// position: per cloud b and axis c, CE between logits x[n, c*NB + k] (n in cloud, k < NB) and soft targets
// prob[90*off[b] + c*(n_b*NB) + n*NB + k]; loss_pos = mean_b mean_c ( lse*sum(p) - sum(p*x) ).
// Writes dlogits = s_pos * (softmax*sum(p) - p) / (3*B) when dlogits != null, with
// s_pos = gout[0] + gout[3]*pos_w (gout = upstream gradient of {pos, rot, open, total}; null -> {0,0,0,1}).
// Trajectory form (motion planner, motion_planner_ptv3.py:326-337): T steps -> 3*T axes per cloud (axis a = t*3 + c,
// logits row width 3*T*NB), row weight mask[b,t] / (3 * sum_t mask[b,t] * B); `total_idx` = index of the total in gout.
// A thread-block cluster of POS_CE_CLUSTER CTAs shares one (cloud, axis): each CTA reduces a slice of the n_b*NB logits, the
// partial (max, sum-exp, sum p, sum p*x) are combined through distributed shared memory, and every CTA writes the gradient
// of its own slice (one (cloud, axis) per CTA left 96 CTAs on 148 SMs for the production batch).
constexpr int POS_CE_CLUSTER = 8;
__global__ void __launch_bounds__(256) k_pos_ce(const float* __restrict__ logits, const float* __restrict__ prob,
                         const int32_t* __restrict__ off, int B, int NB, float* __restrict__ loss_accum,
                         float* __restrict__ dlogits, const float* __restrict__ gout, float pos_w, int T,
                         const float* __restrict__ mask, int total_idx) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), CS = (int)cluster.num_blocks();
  float gscale = gout ? gout[0] + gout[total_idx] * pos_w : pos_w;
  __shared__ float red[32];
  __shared__ float part[4];
  int b = blockIdx.y, c = blockIdx.z;
  int s = off[b], n = off[b + 1] - s;
  int ld = 3 * T * NB;
  int tot = n * NB;
  float wrow = 1.f / (3.f * B);
  if (mask) {
    float len = 0.f;
    for (int t = 0; t < T; ++t) len += mask[b * T + t];
    wrow = mask[b * T + c / 3] / (3.f * len) / B;
  }
  const float* pb = prob + (int64_t)ld * s + (int64_t)c * tot;
  const int chunk = (tot + CS - 1) / CS;
  const int t0 = min(tot, rank * chunk), t1 = min(tot, t0 + chunk);
  // element t of the (cloud, axis) slice is logits[(s + t / NB) * ld + c * NB + t % NB]: (row, column) are carried along the
  // thread's stride instead of divided out per element (three passes x two run-time divisions per element were most of this kernel)
  const int first = t0 + threadIdx.x;
  const int row_first = first / NB, col_first = first - row_first * NB;
  const int drow = blockDim.x / NB, dcol = blockDim.x - drow * NB;
  const float* lbase = logits + (int64_t)s * ld + c * NB;
  float mx = -INFINITY;
  {
    int row = row_first, col = col_first;
    for (int t = first; t < t1; t += blockDim.x) {
      mx = fmaxf(mx, lbase[(int64_t)row * ld + col]);
      row += drow; col += dcol;
      if (col >= NB) { col -= NB; ++row; }
    }
  }
  mx = block_max(mx, red);
  float se = 0.f, sp = 0.f, spx = 0.f;
  {
    int row = row_first, col = col_first;
    for (int t = first; t < t1; t += blockDim.x) {
      float x = lbase[(int64_t)row * ld + col];
      float p = pb[t];
      se += expf(x - mx); sp += p; spx += p * x;
      row += drow; col += dcol;
      if (col >= NB) { col -= NB; ++row; }
    }
  }
  se = block_sum(se, red); sp = block_sum(sp, red); spx = block_sum(spx, red);
  if (threadIdx.x == 0) { part[0] = mx; part[1] = se; part[2] = sp; part[3] = spx; }
  cluster.sync();
  float M = -INFINITY;
  for (int r = 0; r < CS; ++r) M = fmaxf(M, cluster.map_shared_rank(part, r)[0]);
  se = 0.f; sp = 0.f; spx = 0.f;
  for (int r = 0; r < CS; ++r) {
    const float* rp = cluster.map_shared_rank(part, r);
    if (rp[1] > 0.f) se += rp[1] * expf(rp[0] - M);      // empty slices carry (max -inf, sum 0)
    sp += rp[2]; spx += rp[3];
  }
  cluster.sync();                                          // peers may exit only after everyone has read their partials
  float lse = M + logf(se);
  if (threadIdx.x == 0 && rank == 0) atomicAdd(loss_accum, mask ? (lse * sp - spx) * wrow : (lse * sp - spx) / (3.f * B));
  if (dlogits) {
    float k = mask ? gscale * wrow : gscale / (3.f * B);
    int row = row_first, col = col_first;
    float* dbase = dlogits + (int64_t)s * ld + c * NB;
    for (int t = first; t < t1; t += blockDim.x) {
      const int64_t o = (int64_t)row * ld + col;
      dbase[o] = k * (expf(lbase[o] - lse) * sp - pb[t]);
      row += drow; col += dcol;
      if (col >= NB) { col -= NB; ++row; }
    }
  }
}
static int launch_pos_ce(const float* logits, const float* prob, const int32_t* off, int B, int NB, float* loss_accum, float* dlogits,
                         const float* gout, float pos_w, int T, const float* mask, int total_idx, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(POS_CE_CLUSTER, B, 3 * T);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = POS_CE_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LOTUS_CUDA_CALL(cudaLaunchKernelEx(&cfg, k_pos_ce, logits, prob, off, B, NB, loss_accum, dlogits, gout, pos_w, T, mask, total_idx));
  return 0;
}

// rotation CE over 72 bins x 3 axes (logit index bin*3 + axis) + open BCE-with-logits on the last column.
// act: [B, ld] ; gt: [B, 7] = pos(3), rot bins(3), open(1).  losses[1]=rot, [2]=open.
__global__ void k_rot_open_loss(const float* __restrict__ act, int ld, const float* __restrict__ gt, int gt_ld, int B,
                                int nbins, float* __restrict__ losses, float* __restrict__ dact, float rot_w,
                                const float* __restrict__ gout) {
  __shared__ float red[32];
  float s_rot = gout ? gout[1] + gout[3] * rot_w : rot_w;
  float s_open = gout ? gout[2] + gout[3] : 1.f;
  float lrot = 0.f, lopen = 0.f;
  for (int t = threadIdx.x; t < B * 3; t += blockDim.x) {
    int b = t / 3, a = t % 3;
    const float* row = act + (int64_t)b * ld;
    float mx = -INFINITY;
    for (int k = 0; k < nbins; ++k) mx = fmaxf(mx, row[k * 3 + a]);
    float se = 0.f;
    for (int k = 0; k < nbins; ++k) se += expf(row[k * 3 + a] - mx);
    float lse = mx + logf(se);
    int tgt = (int)gt[(int64_t)b * gt_ld + 3 + a];
    lrot += lse - row[tgt * 3 + a];
    if (dact) {
      float kf = s_rot / (3.f * B);
      for (int k = 0; k < nbins; ++k)
        dact[(int64_t)b * ld + k * 3 + a] = kf * (expf(row[k * 3 + a] - lse) - (k == tgt ? 1.f : 0.f));
    }
  }
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float x = act[(int64_t)b * ld + nbins * 3], y = gt[(int64_t)b * gt_ld + gt_ld - 1];
    lopen += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
    if (dact) dact[(int64_t)b * ld + nbins * 3] = s_open * (1.f / (1.f + expf(-x)) - y) / B;
  }
  lrot = block_sum(lrot, red);
  lopen = block_sum(lopen, red);
  if (threadIdx.x == 0) { losses[1] = lrot / (3.f * B); losses[2] = lopen / B; }
}
__global__ void k_total_loss(float* losses, float pos_w, float rot_w) {
  losses[3] = pos_w * losses[0] + rot_w * losses[1] + losses[2];
}

// losses[4] = {pos, rot, open, total}.  When non-null, d_heatmap [N, 3*NB] and d_action [B, act_ld] receive
// sum_i gout[i] * d(losses[i])/d(logits)  (gout: 4 device floats, null -> gradient of `total`).
LOTUS_API int lotus_policy_loss_f32(const float* heatmap_logits, const float* pos_prob, const int32_t* cloud_off, int B,
                                    int NB, const float* action_logits, int act_ld, int rot_bins, const float* gt_actions,
                                    int gt_ld, float pos_w, float rot_w, float* losses, float* d_heatmap,
                                    float* d_action, const float* gout, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  LOTUS_CHECK_ARG(act_ld == rot_bins * 3 + 1, "lotus_policy_loss_f32: action width %d != %d*3+1", act_ld, rot_bins);
  LOTUS_CUDA_CALL(cudaMemsetAsync(losses, 0, 16, st));
  if (int rc = launch_pos_ce(heatmap_logits, pos_prob, cloud_off, B, NB, losses, d_heatmap, gout, pos_w, 1, nullptr, 3, st)) return rc;
  k_rot_open_loss<<<1, 256, 0, st>>>(action_logits, act_ld, gt_actions, gt_ld, B, rot_bins, losses, d_action, rot_w, gout);
  k_total_loss<<<1, 1, 0, st>>>(losses, pos_w, rot_w);
  LOTUS_CUDA_OK();
  return 0;
}

// ---------------------------------------------------------------- trajectory (motion planner) losses
// act: [B*T, ld = nbins*3 + 2] = rot logits (bin*3 + axis), open logit, stop logit; gt: [B*T, 7]; stop_t, mask: [B*T].
// rot = sum(mask*CE)/sum(mask)/3, open/stop = sum(mask*BCE)/sum(mask)   (motion_planner_ptv3.py:363-389)
// losses[5] = {pos, rot, open, stop, total}
__global__ void k_traj_head_loss(const float* __restrict__ act, int ld, const float* __restrict__ gt, int gt_ld,
                                 const float* __restrict__ stop_t, const float* __restrict__ mask, int BT, int nbins,
                                 float* __restrict__ losses, float* __restrict__ dact, float pos_w, float rot_w,
                                 const float* __restrict__ gout) {
  __shared__ float red[32];
  float s_rot = gout ? gout[1] + gout[4] * rot_w : rot_w;
  float s_open = gout ? gout[2] + gout[4] : 1.f;
  float s_stop = gout ? gout[3] + gout[4] : 1.f;
  float msum = 0.f;
  for (int i = threadIdx.x; i < BT; i += blockDim.x) msum += mask[i];
  msum = block_sum(msum, red);
  __shared__ float s_msum;
  if (threadIdx.x == 0) s_msum = msum;
  __syncthreads();
  msum = s_msum;
  float lrot = 0.f, lopen = 0.f, lstop = 0.f;
  for (int t = threadIdx.x; t < BT * 3; t += blockDim.x) {
    int r = t / 3, a = t % 3;
    const float* row = act + (int64_t)r * ld;
    float mx = -INFINITY;
    for (int k = 0; k < nbins; ++k) mx = fmaxf(mx, row[k * 3 + a]);
    float se = 0.f;
    for (int k = 0; k < nbins; ++k) se += expf(row[k * 3 + a] - mx);
    float lse = mx + logf(se);
    int tgt = (int)gt[(int64_t)r * gt_ld + 3 + a];
    float m = mask[r];
    lrot += m * (lse - row[tgt * 3 + a]);
    if (dact) {
      float kf = s_rot * m / msum / 3.f;
      for (int k = 0; k < nbins; ++k)
        dact[(int64_t)r * ld + k * 3 + a] = kf * (expf(row[k * 3 + a] - lse) - (k == tgt ? 1.f : 0.f));
    }
  }
  for (int r = threadIdx.x; r < BT; r += blockDim.x) {
    float m = mask[r];
    float x = act[(int64_t)r * ld + nbins * 3], y = gt[(int64_t)r * gt_ld + gt_ld - 1];
    lopen += m * (fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x))));
    float xs = act[(int64_t)r * ld + nbins * 3 + 1], ys = stop_t[r];
    lstop += m * (fmaxf(xs, 0.f) - xs * ys + log1pf(expf(-fabsf(xs))));
    if (dact) {
      dact[(int64_t)r * ld + nbins * 3] = s_open * m * (1.f / (1.f + expf(-x)) - y) / msum;
      dact[(int64_t)r * ld + nbins * 3 + 1] = s_stop * m * (1.f / (1.f + expf(-xs)) - ys) / msum;
    }
  }
  lrot = block_sum(lrot, red);
  lopen = block_sum(lopen, red);
  lstop = block_sum(lstop, red);
  if (threadIdx.x == 0) {
    losses[1] = lrot / msum / 3.f; losses[2] = lopen / msum; losses[3] = lstop / msum;
    losses[4] = pos_w * losses[0] + rot_w * losses[1] + losses[2] + losses[3];
  }
}
// heatmap_logits [N, T*3*NB] (column (t*3+c)*NB + k), pos_prob packed per cloud as (T, 3, n_b*NB);
// action_logits [B*T, rot_bins*3+2]; gt_trajs [B*T, 7]; stop_targets / traj_mask [B*T] floats (0/1).
// losses[5] = {pos, rot, open, stop, total}; gradients as in lotus_policy_loss_f32 (gout: 5 floats or null).
LOTUS_API int lotus_traj_loss_f32(const float* heatmap_logits, const float* pos_prob, const int32_t* cloud_off, int B, int T,
                                  int NB, const float* action_logits, int act_ld, int rot_bins, const float* gt_trajs,
                                  int gt_ld, const float* stop_targets, const float* traj_mask, float pos_w, float rot_w,
                                  float* losses, float* d_heatmap, float* d_action, const float* gout, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  LOTUS_CHECK_ARG(act_ld == rot_bins * 3 + 2, "lotus_traj_loss_f32: action width %d != %d*3+2", act_ld, rot_bins);
  LOTUS_CHECK_ARG(B > 0 && T > 0, "lotus_traj_loss_f32: empty batch (B=%d, T=%d)", B, T);
  LOTUS_CUDA_CALL(cudaMemsetAsync(losses, 0, 20, st));
  if (int rc = launch_pos_ce(heatmap_logits, pos_prob, cloud_off, B, NB, losses, d_heatmap, gout, pos_w, T, traj_mask, 4, st)) return rc;
  k_traj_head_loss<<<1, 256, 0, st>>>(action_logits, act_ld, gt_trajs, gt_ld, stop_targets, traj_mask, B * T, rot_bins, losses,
                                      d_action, pos_w, rot_w, gout);
  LOTUS_CUDA_OK();
  return 0;
}

// ---------------------------------------------------------------- final-action decode
// position: per cloud and axis argmax over (n_b * NB) logits (== argmax of the softmax), first max wins;
// best = (double)xyz[n, c] + (k - NB/2) * bin_size   (numpy promotes to float64, then the reference casts to f32)
// rotation: per cloud argmax over bins per axis -> euler 'xyz' (extrinsic) -> quaternion xyzw, float64.
// out: [B, 8] float64 = pos(3), quat(4), open logit.
__global__ void k_decode(const float* __restrict__ logits, const float* __restrict__ xyz, int xyz_ld,
                         const int32_t* __restrict__ off, int NB, double bin_size, const float* __restrict__ act, int act_ld,
                         int rot_bins, float resolution, const float* __restrict__ gt_pos, int gt_ld,
                         double* __restrict__ out_, int T, int out_ld) {
  __shared__ float rv[32];
  __shared__ int ri[32];
  int b = blockIdx.x, tstep = blockIdx.y;
  int s = off[b], n = off[b + 1] - s, ld = 3 * T * NB, tot = n * NB;
  logits += tstep * 3 * NB;                       // this step's 3*NB columns
  act += (int64_t)(b * T + tstep) * act_ld;       // row (b, t) of the action logits
  if (gt_pos) gt_pos += (int64_t)(b * T + tstep) * gt_ld;
  double* out = out_ + (int64_t)(b * T + tstep) * out_ld;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = 0; c < 3; ++c) {
    if (gt_pos) {  // compute_final_action=False: pred_pos := gt (simple_policy_ptv3.py:276-277)
      if (threadIdx.x == 0) out[c] = (double)gt_pos[c];
      continue;
    }
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int t = threadIdx.x; t < tot; t += blockDim.x) {
      float v = logits[(int64_t)(s + t / NB) * ld + c * NB + t % NB];
      if (v > best) { best = v; bi = t; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    __syncthreads();
    if (lane == 0) { rv[w] = best; ri[w] = bi; }
    __syncthreads();
    if (w == 0) {
      best = lane < nw ? rv[lane] : -INFINITY;
      bi = lane < nw ? ri[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) {
        int pn = bi / NB, k = bi % NB;
        double cand = (double)xyz[(int64_t)(s + pn) * xyz_ld + c] + (double)(k - NB / 2) * bin_size;
        out[c] = (double)(float)cand;  // torch.from_numpy(np.array(...)).float()
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float* row = act;
    double e[3];
    for (int a = 0; a < 3; ++a) {
      float best = -INFINITY; int bk = 0;
      for (int k = 0; k < rot_bins; ++k)
        if (row[k * 3 + a] > best) { best = row[k * 3 + a]; bk = k; }
      e[a] = ((double)bk * (double)resolution - 180.0) * 0.017453292519943295;
    }
    double cx = cos(e[0] / 2), sx = sin(e[0] / 2), cy = cos(e[1] / 2), sy = sin(e[1] / 2), cz = cos(e[2] / 2), sz = sin(e[2] / 2);
    out[3] = sx * cy * cz - cx * sy * sz;
    out[4] = cx * sy * cz + sx * cy * sz;
    out[5] = cx * cy * sz - sx * sy * cz;
    out[6] = cx * cy * cz + sx * sy * sz;
    out[7] = (double)row[rot_bins * 3];
    if (out_ld > 8) out[8] = (double)row[rot_bins * 3 + 1];   // planner: stop logit
  }
}
LOTUS_API int lotus_decode_actions(const float* heatmap_logits, const float* xyz, int xyz_ld, const int32_t* cloud_off,
                                   int B, int NB, double bin_size, const float* action_logits, int act_ld, int rot_bins,
                                   float euler_resolution, const float* gt_pos_or_null, int gt_ld, double* out,
                                   void* stream_) {
  if (B <= 0) return 0;
  k_decode<<<B, 256, 0, (cudaStream_t)stream_>>>(heatmap_logits, xyz, xyz_ld, cloud_off, NB, bin_size, action_logits, act_ld,
                                                 rot_bins, euler_resolution, gt_pos_or_null, gt_ld, out, 1, 8);
  LOTUS_CUDA_OK();
  return 0;
}
// Trajectory decode (motion_planner_ptv3.py:250-296): out [B, T, 9] float64 = pos(3), quat(4), open logit, stop logit.
LOTUS_API int lotus_decode_trajs(const float* heatmap_logits, const float* xyz, int xyz_ld, const int32_t* cloud_off, int B,
                                 int T, int NB, double bin_size, const float* action_logits, int act_ld, int rot_bins,
                                 float euler_resolution, const float* gt_pos_or_null, int gt_ld, double* out, void* stream_) {
  if (B <= 0 || T <= 0) return 0;
  k_decode<<<dim3(B, T), 256, 0, (cudaStream_t)stream_>>>(heatmap_logits, xyz, xyz_ld, cloud_off, NB, bin_size, action_logits,
                                                          act_ld, rot_bins, euler_resolution, gt_pos_or_null, gt_ld, out, T, 9);
  LOTUS_CUDA_OK();
  return 0;
}
