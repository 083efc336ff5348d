#!/usr/bin/env python
"""Regenerate include/lotus.h from the `LOTUS_API` definitions in robot-3dlotus_b200/csrc/*.cu.

The .cu definitions are the single source of the C ABI; this script copies each prototype verbatim and
attaches the reference interface it replaces (file:line, relative to /root/reference/genrobo3d/).
"""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "robot-3dlotus_b200", "csrc")

DOC = {
    "lotus_last_error": "Message of the last failing call on this thread (reference signals errors with Python exceptions/asserts).",
    "lotus_version": "ABI version (major*10000 + minor*100 + patch).",
    "lotus_device_arch": "Compute capability of the current device (100 on B200).",
    "lotus_grid_coord": "Replaces Point.serialization grid step: models/PointTransformerV3/model.py:96-103 "
                        "(grid = trunc((coord-min_over_batch)/grid_size), depth = bit_length(max)). mode 0 = true "
                        "division (torch CPU kernel), 1 = multiply by reciprocal (torch CUDA kernel). meta[0]=n, meta[1]=depth.",
    "lotus_grid_coord_n": "lotus_grid_coord with the point count read on the device (shape-static launch for CUDA-graph replay with size buckets): model.py:96-103.",
    "lotus_stage_sanitize": "Graph-replay support (train/train_simple_policy.py:202-243 / evaluation/eval_simple_policy.py:316-321 hot loops; SURVEY §8f row 1): makes the table rows between a stage's actual size and its bucket bound inert.",
    "lotus_batch_ids": "Replaces offset2batch: model.py:37-42.",
    "lotus_serialize": "Replaces serialization.encode x4 + argsort + inverse scatter (+ shuffle_orders row permutation): "
                       "serialization/default.py:10-27, z_order.py:66-101, hilbert.py:91-198, model.py:117-138. "
                       "curve_of_row: 0 z, 1 z-trans, 2 hilbert, 3 hilbert-trans (row r of the outputs uses that curve).",
    "lotus_serialize_dev": "lotus_serialize with the rows' curve ids read from device memory (the shuffle_orders draw stays a run-time input of a captured CUDA graph): model.py:117-138.",
    "lotus_pool_structure": "Replaces the integer half of SerializedPooling.forward: model.py:713-755,766-771 "
                            "(code>>3, unique -> cluster, CSR over the code-sorted order, head attributes, pooled depth/offsets).",
    "lotus_patch_plan": "Replaces SerializedAttention.get_padding_and_inverse and the order[pad] / unpad[inverse] "
                        "compositions: model.py:410-466,480-481. gather_idx = order[pad]; slot_of_point = unpad[inverse].",
    "lotus_neighbor_table": "Replaces spconv 2.3.6 SubMConv3d indice-pair generation (call sites model.py:616-622,845-852): "
                            "nbr[d*stride + p] = voxel at grid[p]+off(d) in the same cloud or -1; d=(d0*k+d1)*k+d2, d0 along x.",
    "lotus_pair_lists": "Compact per-tap (output, input) voxel pair lists — the rulebook of spconv 2.3.6 SubMConv3d (call sites model.py:616-622,845-852); "
                        "consumed by the weight-gradient kernels.",
    "lotus_policy_step": "One whole policy step in a single call: replaces SimplePolicyPTV3CA.forward + compute_loss + loss.backward() "
                         "(models/simple_policy_ptv3.py:225-373; models/PointTransformerV3/model_ca.py:383-412) for the production topology; "
                         "parameter gradients accumulate into the caller's sinks (train/train_simple_policy.py:211-216).",
    "lotus_policy_step_profile": "Measurement aid: switch lotus_policy_step's profile mode (CUDA events around every launcher call, side stream folded back).",
    "lotus_policy_step_profile_count": "Number of launcher calls recorded by the last profiled lotus_policy_step.",
    "lotus_policy_step_profile_read": "One record of the last profiled step: C-ABI name, device milliseconds, the call's arguments (bench.py per-kernel table).",
    "lotus_policy_step_calls": "Kernel-launching calls issued by the last lotus_policy_step (bench accounting).",
    "lotus_attn_debug_buffer": "Development aid: per-CTA clock64 stamps of the fused patch-attention phases.",
    "lotus_tc_debug_buffer": "Development aid: per-CTA %globaltimer stamps of the tensor-core GEMM phases.",
    "lotus_fps": "Replaces farthest_point_sample (explicit start index): train/datasets/common.py:108-129. Bit-exact indices.",
    "lotus_knn": "North-star kNN grouping (no reference function; SURVEY §8a row 13): brute force inside each cloud, ties -> lowest index.",
    "lotus_ball_query": "North-star radius grouping (no reference function; SURVEY §8a row 13): first nsample indices within radius.",
    "lotus_rpe_bias": "Replaces RPE.forward + get_rel_pos: model.py:307-326,400-408 (materialised bias; the attention kernels fuse it).",
    "lotus_linear_fwd_f32": "Replaces nn.Linear forward (qkv/proj/fc1/fc2/q/kv/pool/unpool/txt_fc/head): model.py:386-387,572-574,707,804-805; "
                            "model_ca.py:27-31; simple_policy_ptv3.py:40-68,387. Optional row gather, bias, GELU/LeakyReLU, residual epilogue.",
    "lotus_linear_dgrad_f32": "Backward of nn.Linear w.r.t. its input (autograd of the call sites above).",
    "lotus_linear_wgrad_f32": "Backward of nn.Linear w.r.t. its weight (accumulates into dw).",
    "lotus_subm_conv_fwd_f32": "Replaces spconv.SubMConv3d forward (stem k=5, CPE k=3): model.py:616-622,845-852.",
    "lotus_subm_conv_dgrad_f32": "Backward of SubMConv3d w.r.t. its input features.",
    "lotus_subm_conv_wgrad_f32": "Backward of SubMConv3d w.r.t. its weight (accumulates into dw).",
    "lotus_linear_fwd_bf16": "bf16 tensor-core (tcgen05/TMEM/TMA) form of nn.Linear forward, same call sites as lotus_linear_fwd_f32; "
                             "fused bias/GELU/LeakyReLU/dropout/residual epilogue, optional pre-activation copy.",
    "lotus_linear_dgrad_bf16": "bf16 tensor-core backward of nn.Linear w.r.t. its input (uses the pre-transposed bf16 weight).",
    "lotus_linear_dgrad_bias_bf16": "lotus_linear_dgrad_bf16 that also accumulates the layer's bias gradient (column sums of dy, read from the operand tiles already in shared memory): "
                                    "autograd of nn.Linear at the call sites of lotus_linear_fwd_f32.",
    "lotus_linear_dgrad_act_bias_bf16": "lotus_linear_dgrad_act_bf16 that also accumulates fc2's bias gradient (column sums of dy): model.py:560-583.",
    "lotus_linear_dgrad_act_bf16": "bf16 tensor-core backward of `fc2(drop(GELU(u)))` w.r.t. u (MLP, model.py:560-583): data-gradient GEMM with gelu'(u) and the forward's dropout mask applied in the epilogue.",
    "lotus_disc_pos_prob": "Replaces get_disc_gt_pos_prob (soft position labels, computed per keystep by the DataLoader workers): utils/action_position_utils.py:7-46, "
                           "call site train/datasets/simple_policy_dataset.py:346-354. float64 arithmetic like numpy, packed fp32 output.",
    "lotus_subsample_keysteps": "Replaces the per-keystep point subsampling of the DataLoader workers (np.random.permutation(len)[:k] / np.random.choice(len, k, replace=False)): "
                                "train/datasets/simple_policy_dataset.py:281-294. Uniform random k-subset in random order from counter-based keys sorted per cloud.",
    "lotus_augment_keysteps": "Replaces the per-keystep geometric augmentation + feature assembly of the DataLoader workers: height, z-rotation, U(0, 2 mm) jitter, centring, "
                              "colour scaling: train/datasets/simple_policy_dataset.py:158-181,296,321-346; train/datasets/common.py:81-87. float64 arithmetic like numpy.",
    "lotus_voxel_downsample": "Replaces the workspace crop + open3d voxel_down_sample_and_trace of Actioner.process_point_clouds: evaluation/eval_simple_policy.py:151-175 "
                              "(voxel mean, colour of the voxel's first input point; voxels ordered by first appearance).",
    "lotus_voxel_downsample_scratch_bytes": "Scratch size (bytes) lotus_voxel_downsample needs for n input points.",
    "lotus_voxel_bounds": "The crop box and open3d's voxel_min_bound (min of the cropped cloud - voxel_size / 2) on the device: evaluation/eval_simple_policy.py:154-173.",
    "lotus_disc_pos_prob_scratch_bytes": "Scratch size (bytes) lotus_disc_pos_prob needs for B clouds.",
    "lotus_linear_wgrad_bf16": "bf16 tensor-core backward of nn.Linear w.r.t. its weight: split-K, MN-major operands read in place, fp32 accumulate.",
    "lotus_subm_conv_fwd_bf16": "bf16 tensor-core implicit-GEMM form of spconv.SubMConv3d forward (CPE k=3): model.py:616-622; gathered operand staged by cp.async.",
    "lotus_subm_conv_dgrad_bf16": "bf16 tensor-core backward of SubMConv3d w.r.t. its input features (mirrored taps).",
    "lotus_subm_conv_wgrad_bf16": "bf16 tensor-core backward of SubMConv3d w.r.t. its weight (per-tap split-K, fp32 accumulate).",
    "lotus_subm_conv_wgrad_pairs_bf16": "bf16 tensor-core backward of SubMConv3d w.r.t. its weight over the compact rulebook (lotus_pair_lists): "
                                        "one launch for all 27 taps, both operands gathered by cp.async, fp32 accumulate.",
    "lotus_layernorm_fwd_bf16": "bf16-activation form of lotus_layernorm_fwd_f32 (same reference call sites).",
    "lotus_layernorm_bwd_bf16": "bf16-activation form of lotus_layernorm_bwd_f32 (dy may be fp32: gradients leaving the attention kernels).",
    "lotus_colsum_bf16": "bf16-activation form of lotus_colsum_f32 (bias gradients).",
    "lotus_batchnorm_fwd_bf16": "bf16-activation form of lotus_batchnorm_fwd_f32 (the stem feeds it fp32): model.py:932,785-788,854-857.",
    "lotus_batchnorm_bwd_bf16": "bf16-activation form of lotus_batchnorm_bwd_f32.",
    "lotus_subm_conv_splits": "How many reduction ranges the split-K conv forms give a shape (1: use the plain entry points).",
    "lotus_subm_conv_fwd_splitk_bf16": "lotus_subm_conv_fwd_bf16 with the 27 * Cin reduction split over the grid (deep stages: few rows, long K); partial tiles meet in an fp32 scratch.",
    "lotus_subm_conv_dgrad_splitk_bf16": "Split-K form of lotus_subm_conv_dgrad_bf16.",
    "lotus_linear_wgrad_bias_bf16": "nn.Linear weight AND bias gradient in one tcgen05 GEMM (db = dy^T 1 through one extra MMA per K step against a tile of ones): model.py:386-387,572-574.",
    "lotus_layernorm_bwd_mask_bf16": "lotus_layernorm_bwd_bf16 (contiguous rows) that also writes dropout_mask(dx) for the next layer's output dropout (nn.Dropout backward of proj / fc2: model.py:555-557,577-583).",
    "lotus_batchnorm_bwd_acc_bf16": "Training-mode backward of BatchNorm1d + GELU in ONE cooperative launch (sums, grid barrier, dx) that also adds the batch sums into the parameter-gradient accumulators: model.py:932,785-788,854-857.",
    "lotus_pointwise_bf16": "bf16-activation form of lotus_pointwise_f32 (GELU / LeakyReLU and derivatives with fused dropout).",
    "lotus_convert": "fp32 <-> bf16 conversion between the fp32 head/loss tensors and the bf16 trunk.",
    "lotus_convert_rows_f32_to_bf16": "strided fp32 -> bf16 row copy (attention dV into the packed dqkv gradient): model.py:541 (torch.stack of q,k,v) in reverse.",
    "lotus_cast_transpose_bf16": "stages fp32 master weights for the tensor-core path: bf16 copy and [B][D][A] transpose (nn.Linear / SubMConv3d weights).",
    "lotus_cast_transpose_batch_bf16": "All weights of the model staged for the tensor-core kernels in one launch (host arrays of per-weight pointers and extents); same per-weight semantics as lotus_cast_transpose_bf16.",
    "lotus_dropout_bf16": "bf16-activation form of lotus_dropout_f32.",
    "lotus_patch_attn_fused_fwd_bf16": "Replaces SerializedAttention.forward between the qkv and proj linears in ONE tcgen05 kernel: qkv[order] gather, per-head q/k LayerNorm "
                                       "(model.py:393-394,532-535), QK^T, scale + RPE bias (model.py:307-326,400-408,520-521), softmax, attention dropout, PV, feat[inverse] scatter: model.py:497-551.",
    "lotus_patch_attn_fused_bwd_bf16": "Backward of lotus_patch_attn_fused_fwd_bf16: bf16 gradients w.r.t. the raw q/k/v rows written straight into the packed dqkv tensor "
                                       "(per-head LayerNorm backward in the epilogue), q/k-norm affine and RPE table gradients accumulated.",
    "lotus_patch_attn_fwd_bf16": "bf16-activation form of lotus_patch_attn_fwd_f32: model.py:543-551.",
    "lotus_patch_attn_bwd_bf16": "bf16-activation form of lotus_patch_attn_bwd_f32 (fp32 gradient accumulators).",
    "lotus_cross_attn_fwd_bf16": "bf16-activation form of lotus_cross_attn_fwd_f32: model_ca.py:56-67.",
    "lotus_cross_attn_bwd_bf16": "bf16-activation form of lotus_cross_attn_bwd_f32 (fp32 gradient outputs).",
    "lotus_segment_max_fwd_bf16": "bf16-activation form of lotus_segment_max_fwd_f32: model.py:760-762.",
    "lotus_segment_max_bwd_bf16": "bf16-activation form of lotus_segment_max_bwd_f32.",
    "lotus_segment_sum_bf16": "bf16-activation form of lotus_segment_sum_f32.",
    "lotus_unpool_add_bf16": "bf16-activation form of lotus_unpool_add_f32: model.py:824.",
    "lotus_layernorm_fwd_f32": "Replaces nn.LayerNorm (block norms, CPE norm, per-head q/k norm eps 1e-6): model.py:393-394,624,627,645; model_ca.py:43-44,114,124. "
                               "Rows are addressed as (r/inner)*outer_stride + (r%inner)*inner_stride so per-head norms read the packed qkv in place.",
    "lotus_layernorm_bwd_f32": "Backward of nn.LayerNorm (dx, and accumulated dgamma/dbeta).",
    "lotus_colsum_f32": "Bias gradients (column sums of dy), accumulated.",
    "lotus_batchnorm_fwd_f32": "Replaces nn.BatchNorm1d(eps=1e-3, momentum=0.01) [+ GELU] (stem, pool, unpool): model.py:932,785-788,807-813,854-857.",
    "lotus_batchnorm_bwd_f32": "Backward of BatchNorm1d [+ GELU].",
    "lotus_pointwise_f32": "GELU / LeakyReLU(0.02) / SiLU and their derivatives, add, scale: model.py:573,279; simple_policy_ptv3.py:42,65.",
    "lotus_dropout_f32": "Replaces nn.Dropout (counter-based mask; the same call on dy is the backward): model.py:388,575; simple_policy_ptv3.py:43,66.",
    "lotus_patch_attn_fwd_f32": "Replaces flash_attn_varlen_qkvpacked_func over the serialized patches incl. the qkv[order] gather and feat[inverse] scatter "
                                "(and the RPE bias of the non-flash branch): model.py:497,520-521,543-551.",
    "lotus_patch_attn_bwd_f32": "Backward of the patch attention (dq, dk, dv accumulated in point order; optional RPE table gradient).",
    "lotus_cross_attn_fwd_f32": "Replaces flash_attn_varlen_kvpacked_func (points of a cloud attend to that cloud's tokens): model_ca.py:56-67.",
    "lotus_cross_attn_bwd_f32": "Backward of the cross attention.",
    "lotus_segment_max_fwd_f32": "Replaces torch_scatter.segment_csr(reduce='max') and the per-cloud max of the action head: model.py:760-762; simple_policy_ptv3.py:117-119.",
    "lotus_segment_max_bwd_f32": "Backward of the segment max (routes dy to the arg-max rows).",
    "lotus_segment_sum_f32": "Replaces torch_scatter.segment_csr(reduce='mean') for coords (model.py:763-765); sum mode is the backward of the unpool gather.",
    "lotus_unpool_add_f32": "Replaces parent.feat + point.feat[inverse]: model.py:824.",
    "lotus_policy_loss_f32": "Replaces compute_loss (heatmap_disc soft-label CE, euler_disc CE, open BCE) and its backward: simple_policy_ptv3.py:308-373.",
    "lotus_traj_loss_f32": "Replaces MotionPlannerPTV3*.compute_loss (masked trajectory soft-label CE, rot CE, open/stop BCE) and its backward: motion_planner_ptv3.py:307-397.",
    "lotus_decode_trajs": "Replaces the trajectory decode (T steps per keystep, stop logit appended): motion_planner_ptv3.py:250-296.",
    "lotus_pdnorm_modulate_fwd_f32": "Replaces the adaptive modulation of PDNorm, feat*(1+scale)+shift with per-cloud [shift|scale]: model.py:293-303.",
    "lotus_pdnorm_modulate_bwd_f32": "Backward of the PDNorm modulation (dx and the per-cloud d[shift|scale]).",
    "lotus_decode_actions": "Replaces the final-action decode: simple_policy_ptv3.py:250-296; utils/action_position_utils.py:48-60; utils/rotation_transform.py:192-193.",
    "lotus_grad_sumsq_f32": "Global gradient norm of clip_grad_norm_: train/train_simple_policy.py:237-241.",
    "lotus_adamw_step_f32": "Replaces the reference AdamW.step over one flat buffer (+ fused clip coefficient): train/optim/adamw.py:53-112.",
}


def prototypes():
    out = []
    for f in sorted(glob.glob(os.path.join(SRC, "*.cu"))):
        src = open(f).read()
        for m in re.finditer(r"LOTUS_API\s+([^\{;]+?)\s*\{", src, re.S):
            proto = " ".join(m.group(1).split())
            name = re.search(r"(lotus_\w+)\s*\(", proto).group(1)
            out.append((name, proto, os.path.basename(f)))
    return out


def main():
    lines = [
        "/* lotus.h — C ABI of the B200-native 3D-LOTUS hot path (generated by tools/gen_header.py; do not edit).",
        " *",
        " * Conventions: every function returns 0 on success, non-zero on failure (message: lotus_last_error());",
        " * pointers are DEVICE pointers unless the parameter name says host; the library never allocates —",
        " * callers pass outputs and scratch; `stream` is a cudaStream_t; calls are stream-ordered and re-entrant",
        " * per stream.  Index arrays are int32, curve codes int64.  Citations are relative to",
        " * /root/reference/genrobo3d/ (models/PointTransformerV3/ for bare model.py / model_ca.py / serialization/).",
        " */",
        "#ifndef LOTUS_H_",
        "#define LOTUS_H_",
        "#include <stdint.h>",
        "#ifdef __cplusplus",
        'extern "C" {',
        "#endif",
        "",
    ]
    missing = []
    for name, proto, f in prototypes():
        doc = DOC.get(name)
        if doc is None:
            missing.append(name)
            doc = "(undocumented)"
        lines.append("/* [%s] %s */" % (f, doc))
        lines.append(proto + ";")
        lines.append("")
    lines += ["#ifdef __cplusplus", "}", "#endif", "#endif  /* LOTUS_H_ */", ""]
    open(os.path.join(ROOT, "include", "lotus.h"), "w").write("\n".join(lines))
    if missing:
        print("WARNING: undocumented:", missing)
    print("wrote include/lotus.h with", len(prototypes()), "entry points")


if __name__ == "__main__":
    main()
