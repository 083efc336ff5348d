"""Training-loop shell: command-line config merge, schedules, checkpoint layout, resume."""
import importlib.util
import math
import os

import pytest
import torch

from conftest import pkg

REF_SCHED = "/root/reference/genrobo3d/train/optim/sched.py"


def test_cli_merge_order_and_literals(tmp_path):
    tl = pkg("train_loop")
    y = tmp_path / "a.yaml"
    y.write_text("SEED: 2024\noutput_dir: out\nTRAIN:\n  train_batch_size: 64\n  lr_sched: cosine\nMODEL:\n  model_class: SimplePolicyPTV3AdaNorm\n"
                 "  ptv3_config:\n    enc_channels: [32, 64]\n    qk_norm: false\n")
    args, cfg = tl.build_args(["--exp-config", str(y), "TRAIN.train_batch_size", "8", "MODEL.model_class", "SimplePolicyPTV3CA",
                               "MODEL.ptv3_config.enc_channels", "[64,128,256,512,768]", "MODEL.ptv3_config.qk_norm", "True",
                               "TRAIN.learning_rate", "1e-4", "MODEL.action_config.pos_pred_type", "heatmap_disc"])
    assert cfg.SEED == 2024 and cfg.TRAIN.train_batch_size == 8 and cfg.TRAIN.warmup_steps == 5000       # default <- file <- opts
    assert cfg.MODEL.model_class == "SimplePolicyPTV3CA" and cfg.MODEL.ptv3_config.enc_channels == [64, 128, 256, 512, 768]
    assert cfg.MODEL.ptv3_config.qk_norm is True and cfg.TRAIN.learning_rate == 1e-4
    assert cfg.MODEL.action_config.pos_pred_type == "heatmap_disc" and cfg.is_frozen()
    with pytest.raises(ValueError):
        tl.get_config(str(y), ["TRAIN.train_batch_size"])


@pytest.mark.parametrize("sched", ["linear", "inverse_sqrt", "cosine", "cosine_cycle"])
def test_schedules_match_reference(sched):
    tl, cfgm = pkg("train_loop"), pkg("config")
    T = cfgm.ConfigNode(lr_sched=sched, warmup_steps=50, num_train_steps=1000, num_cosine_cycles=3, learning_rate=1e-4)
    steps = [0, 1, 25, 49, 50, 51, 333, 500, 999, 1000]
    ours = [tl.lr_decay_rate(s, T) for s in steps]
    assert ours[0] == 1e-5 and abs(ours[2] - 0.5) < 1e-12                     # warm-up is linear, floored at 1e-5
    if sched == "cosine":
        assert abs(ours[-1] - 1e-5) < 1e-12 and abs(ours[4] - 1.0) < 1e-12
        assert abs(ours[7] - 0.5 * (1 + math.cos(math.pi * (500 - 50) / 950))) < 1e-12
    if os.path.exists(REF_SCHED):                                             # the reference's own schedule code, when mounted
        spec = importlib.util.spec_from_file_location("ref_sched", REF_SCHED)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        assert ours == [ref.get_lr_sched_decay_rate(s, T) for s in steps]


@pytest.mark.gpu
def test_train_save_resume_roundtrip(tmp_path):
    """3 optimizer steps -> checkpoint files in the reference layout -> resume continues from the saved step with identical
    parameters and Adam state; gradient accumulation of two half-batches equals one step on the whole batch"""
    tl, cfgm, syn = pkg("train_loop"), pkg("config"), pkg("synthetic")
    out = str(tmp_path / "exp")
    model_cfg = cfgm.small_model_config(attn_drop=0.0, proj_drop=0.0, enc_channels=[64, 64, 64, 64, 128], enc_num_head=[2, 2, 2, 2, 4],
                                        dec_channels=[64, 64, 64, 64], dec_num_head=[2, 2, 2, 2])
    model_cfg.action_config.dropout = 0.0
    cfg = tl.get_config(None, ["output_dir", out, "SEED", "7", "TRAIN.num_train_steps", "3", "TRAIN.save_steps", "2", "TRAIN.log_steps", "1",
                               "TRAIN.warmup_steps", "2", "TRAIN.grad_norm", "10"])
    cfg.defrost()
    cfg.MODEL = model_cfg
    cfg.freeze()
    batches = [syn.make_batch(seed=50 + i, sizes=[300, 200]) for i in range(3)]
    dev = torch.device("cuda", 0)
    model, step_fn = tl.train(cfg, lambda e: iter(batches), 3, val_batches=batches[:1], device=dev)
    ck = os.path.join(out, "ckpts")
    assert sorted(os.listdir(ck)) == ["model_step_2.pt", "model_step_3.pt", "train_state_latest.pt"]
    st = torch.load(os.path.join(ck, "train_state_latest.pt"), weights_only=False)
    assert st["step"] == 3 and set(st["optimizer"]) == {"state", "param_groups"}
    sd = torch.load(os.path.join(ck, "model_step_3.pt"), weights_only=False)
    assert all(v.device.type == "cpu" for v in sd.values() if isinstance(v, torch.Tensor))
    assert os.path.exists(os.path.join(out, "logs", "training_config.yaml")) and os.path.exists(os.path.join(out, "logs", "log.txt"))
    # resume: nothing left to do at step 3 -> the restored state equals the saved one
    cfg2 = tl.get_config(None, ["output_dir", out, "SEED", "7", "TRAIN.num_train_steps", "3", "TRAIN.save_steps", "2", "TRAIN.warmup_steps", "2",
                                "TRAIN.grad_norm", "10"])
    cfg2.defrost()
    cfg2.MODEL = model_cfg
    cfg2.freeze()
    model2, step2 = tl.train(cfg2, lambda e: iter(batches), 3, device=dev)
    assert step2.global_step == 3
    for (n1, p1), (n2, p2) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert n1 == n2 and torch.equal(p1.cpu(), p2.cpu()), n1
    assert torch.equal(step_fn.opt.exp_avg.cpu(), step2.opt.exp_avg.cpu())


def test_lookahead_pairs():
    tl = pkg("train_loop")
    assert list(tl._lookahead([])) == []
    assert list(tl._lookahead([1])) == [(1, None)]
    assert list(tl._lookahead(iter([1, 2, 3]))) == [(1, 2), (2, 3), (3, None)]


def test_model_saver_layout(tmp_path):
    """ModelSaver writes `model_step_N.pt` (CPU tensors, 'module.' stripped) and `train_state_latest.pt` = {'step', 'optimizer'}
    exactly where the reference's resume logic looks for them (train/utils/save.py:20-45, train_simple_policy.py:130-146)"""
    tl = pkg("train_loop")

    class Opt:
        def state_dict(self):
            return {"state": {}, "param_groups": []}

    class Wrapped(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.module = torch.nn.Linear(3, 2)

    saver = tl.ModelSaver(str(tmp_path))
    saver.save(Wrapped(), 40, optimizer=Opt(), rewrite_optimizer=True)
    saver.save(Wrapped(), 50, optimizer=Opt(), rewrite_optimizer=False)
    assert sorted(os.listdir(tmp_path)) == ["model_step_40.pt", "model_step_50.pt", "train_state_50.pt", "train_state_latest.pt"]
    sd = torch.load(tmp_path / "model_step_40.pt")
    assert sorted(sd) == ["bias", "weight"]
    assert torch.load(tmp_path / "train_state_latest.pt", weights_only=False)["step"] == 40
