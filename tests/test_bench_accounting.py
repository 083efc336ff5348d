"""bench.py's byte / FLOP model reads call arguments BY POSITION: pin those positions against the parameter names in
include/lotus.h (a shifted index once turned the conv data-gradient TFLOP/s into nonsense and crashed the LayerNorm rows)."""
import importlib.util
import os

from conftest import ROOT, pkg


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _names(protos, fn):
    return [n for _, n in protos[fn][1]]


def test_positional_indices_match_header():
    b = _bench()
    protos = pkg("lib").parse_header()
    for fn, (i, j, k, _) in b.GEMM_ARGS.items():
        n = _names(protos, fn)
        assert (n[i], n[j], n[k]) == ("M", "N", "K"), (fn, n)
    for fn, i in b.CONV_ARGS.items():
        n = _names(protos, fn)
        assert n[i] in ("n", "n_upper") and n[i + 1: i + 4] == ["Cin", "Cout", "K3"], (fn, n)
    for fn in ("lotus_patch_attn_fwd_bf16", "lotus_patch_attn_bwd_bf16"):
        assert _names(protos, fn)[9:13] == ["n_patches", "max_len", "H", "Dh"]
    for fn in ("lotus_cross_attn_fwd_bf16", "lotus_cross_attn_bwd_bf16"):
        n = _names(protos, fn)
        assert n[6] == "pt_off" and n[8] == "B" and n[11:13] == ["H", "Dh"]
    n = _names(protos, "lotus_layernorm_fwd_bf16")
    assert n[8] == "residual" and n[13:15] == ["rows", "C"]
    n = _names(protos, "lotus_layernorm_bwd_bf16")
    assert n[1] == "dy_fp32" and n[18:21] == ["rows", "C", "accumulate_dx"]
    n = _names(protos, "lotus_layernorm_bwd_f32")
    assert n[17:20] == ["rows", "C", "accumulate_dx"]
    assert _names(protos, "lotus_pointwise_bf16")[1:4] == ["a", "y", "n"]
    assert _names(protos, "lotus_dropout_bf16")[2] == "n"
    assert _names(protos, "lotus_colsum_bf16")[1:3] == ["rows", "C"]
    assert _names(protos, "lotus_batchnorm_fwd_bf16")[1:5] == ["x_fp32", "y", "rows", "C"]
    assert _names(protos, "lotus_batchnorm_bwd_bf16")[2:5] == ["x_fp32", "rows", "C"]
    assert _names(protos, "lotus_adamw_step_f32")[4] == "n" and _names(protos, "lotus_grad_sumsq_f32")[1] == "n"
    assert _names(protos, "lotus_convert_rows_f32_to_bf16")[4:6] == ["rows", "C"]
    assert _names(protos, "lotus_linear_fwd_bf16")[5] == "y_preact"


def test_algorithmic_work_formulas():
    b = _bench()
    ctx = {"conv_pairs": {(1000, 27): 12000}, "n_of_upper": {}, "npad_of_patches": {8: 1024}, "n_of_cloud_off": {77: 1000}, "Lt": 64}
    fl, by = b.algorithmic_work("lotus_linear_fwd_bf16", (1, 1, 1, 0, 1, 0, 1000, 256, 64, 0, 0, 0, 0, 0, 0), ctx)
    assert fl == 2.0 * 1000 * 256 * 64 and by == 2.0 * (1000 * 64 + 256 * 64 + 1000 * 256)
    fl2, by2 = b.algorithmic_work("lotus_linear_fwd_bf16", (1, 1, 1, 0, 1, 5, 1000, 256, 64, 1, 0, 1, 0, 0, 0), ctx)
    assert fl2 == fl and by2 == by + 2.0 * 1000 * 256                         # second (pre-activation) output
    fl, by = b.algorithmic_work("lotus_subm_conv_dgrad_bf16", (1, 1, 1, 1, 1, 5000, 1000, 64, 64, 27, 0, 0), ctx)
    assert fl == 2.0 * 12000 * 64 * 64                                        # the ACTUAL pair count, not n * 27
    fl, by = b.algorithmic_work("lotus_patch_attn_fwd_bf16", (1, 1, 1, 64, 64, 192, 1, 1, 1, 8, 128, 2, 32) + (0,) * 11, ctx)
    assert by == 2.0 * 4 * 1024 * 64 + 8.0 * 1024 + 4.0 * 1024 * 2
    fl, by = b.algorithmic_work("lotus_cross_attn_fwd_bf16", (1, 64, 1, 1, 64, 128, 77, 5, 4, 400, 30, 2, 32) + (0,) * 8, ctx)
    assert by == 2.0 * (2 * 1000 * 64 + 2 * 64 * 64) + 4.0 * 1000 * 2
    fl, by = b.algorithmic_work("lotus_layernorm_bwd_bf16", (1, 0, 1, 64, 0, 1, 1, 64, 0, 1, 1, 64, 0, 1, 1, 1, 1, 1, 1000, 64, 1, 0), ctx)
    assert fl == 0 and by == 1000 * 64 * (2 + 2 + 4) + 8.0 * 1000
    assert b.algorithmic_work("lotus_adamw_step_f32", (1, 1, 1, 1, 1000) + (0,) * 13, ctx) == (0.0, 28000.0)


def test_traffic_db_covers_every_roofline_kernel():
    """`roofline.traffic` / `attention_block.traffic` come from profiles/ncu_traffic.json: every device kernel bench.py can name as dominant
    (KERNEL_OF values that the round-2 step actually launches) and the attention pair must have a measured DRAM-bytes entry with its source"""
    import json
    b = _bench()
    db = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    need = {"lin::k_lin", "tc::k_gemm_tc", "paf::k_pattn_fwd", "paf::k_pattn_bwd", "patch_attention(fwd+bwd)"}
    assert need <= set(db), sorted(need - set(db))
    assert need - {"patch_attention(fwd+bwd)"} <= set(b.KERNEL_OF.values())
    for k in need:
        assert db[k]["dram_bytes_per_launch"] > 0 and db[k]["launches"] > 0 and "ncu" in db[k]["source"]
