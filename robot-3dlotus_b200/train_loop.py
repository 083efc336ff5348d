"""Training-loop shell with the reference's command line, schedule and checkpoint layout, driving the
B200 engine (`trainer.TrainStep`: native forward/backward, ONE gradient all-reduce, fused clip + AdamW).

Reference behaviour kept (file:line under /root/reference/genrobo3d/):
  * CLI `--exp-config a.yaml[;b.yaml] KEY VAL KEY VAL ...`                  train/train_simple_policy.py:314-343,
    merge order defaults <- files <- trailing options                       configs/default.py:62-94 (yacs `merge_from_list`:
    dotted keys, values parsed as Python literals, lists from `a,b,c`)
  * per-rank seed = SEED + rank                                            train/train_simple_policy.py:64-67
  * resume: `ckpts/train_state_latest.pt` = {'step', 'optimizer'} ->
    `ckpts/model_step_<step>.pt`, else `config.checkpoint` (shape-filtered,
    `checkpoint_strict_load`)                                              :130-173
  * lr = init_lr * max(decay(global_step), 1e-5) floor 1e-8, warm-up +
    cosine / linear / inverse_sqrt / cosine_cycle                          :218-221, train/optim/sched.py:9-113
  * gradient accumulation, grad-norm clip, optimizer step, zero_grad       :205-232
  * `model_step_N.pt` (CPU state dict, 'module.' stripped) + optimizer
    state every `save_steps` and at the end                                train/utils/save.py:20-45
  * validation metrics total/pos/rot/open loss, pos L1, open accuracy      :271-300
  * logs/training_config.yaml, logs/log.txt                                 train/utils/save.py:13-18

The batch producer (LMDB + msgpack dataset) is not part of this path: `train()` takes any iterable
factory yielding `ptv3_collate_fn`-shaped batches (Appendix B); the CLI offers `--synthetic` (synthetic.make_batch)
so the shell runs end to end offline.  SyncBatchNorm is not applied (per-rank statistics).
"""
import argparse
import ast
import logging
import math
import os
import sys
import time

import torch
import yaml

from .config import ConfigNode
from .policy import MODEL_FACTORY
from .trainer import TrainStep, warmup_cosine

LOGGER = logging.getLogger("lotus.train")

# configs/default.py:17-48
_DEFAULTS = dict(
    SEED=2023, output_dir=None, checkpoint=None, checkpoint_strict_load=True, tfboard_log_dir=None,
    TRAIN=dict(resume_training=True, train_batch_size=8, gradient_accumulation_steps=1, num_epochs=None, num_train_steps=100000,
               warmup_steps=5000, log_steps=1000, save_steps=5000, val_steps=None, optim="adamw", learning_rate=1e-4,
               lr_sched="cosine", num_cosine_cycles=None, betas=[0.9, 0.98], weight_decay=0.05, grad_norm=5, n_workers=0,
               pin_mem=True),
    MODEL=dict(), DATASET=dict(),
)


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = ConfigNode(v) if isinstance(v, dict) else v


def _literal(v):
    """yacs `_decode_cfg_value`: strings that parse as Python literals become those literals, everything else stays a string"""
    if not isinstance(v, str):
        return v
    try:
        return ast.literal_eval(v)
    except (ValueError, SyntaxError):
        return v


def get_config(config_paths=None, opts=None):
    """configs/default.py:62-94"""
    cfg = ConfigNode(_DEFAULTS)
    if config_paths:
        paths = config_paths.split(";") if isinstance(config_paths, str) else list(config_paths)
        for p in paths:
            with open(p) as f:
                _merge(cfg, yaml.safe_load(f) or {})
    opts = list(opts or [])
    if len(opts) % 2:
        raise ValueError("override list has odd length: %r; it must be a list of KEY VALUE pairs" % (opts,))
    for key, val in zip(opts[0::2], opts[1::2]):
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            if p not in node:
                node[p] = ConfigNode()
            node = node[p]
        node[parts[-1]] = _literal(val)
    cfg.freeze()
    return cfg


def lr_decay_rate(global_step, T):
    """train/optim/sched.py:95-113 (and the schedule functions at :16-70)"""
    sched, warm, tot = T.lr_sched, T.warmup_steps, T.num_train_steps
    if global_step < warm:
        r = global_step / warm
    elif sched == "linear":
        r = max(0, (tot - global_step) / (tot - warm))
    elif sched == "inverse_sqrt":
        r = warm ** 0.5 * global_step ** -0.5
    elif sched == "cosine":
        r = warmup_cosine(global_step, warm, tot)
    elif sched == "cosine_cycle":                       # hard restarts (sched.py:56-63)
        progress = float(global_step - warm) / float(max(1, tot - warm))
        r = 0.0 if progress >= 1.0 else 0.5 * (1.0 + math.cos(math.pi * ((float(T.num_cosine_cycles) * progress) % 1.0)))
    elif sched == "stepwise":
        r = T.lr_decay_gamma ** (global_step // T.lr_decay_step_size)
    else:
        raise NotImplementedError(f"invalid lr scheduler {sched}")
    return max(r, 1e-5)


class ModelSaver:
    """train/utils/save.py:20-45"""

    def __init__(self, output_dir, prefix="model_step", suffix="pt"):
        self.output_dir, self.prefix, self.suffix = output_dir, prefix, suffix

    def save(self, model, step, optimizer=None, rewrite_optimizer=False):
        sd = {}
        for k, v in model.state_dict().items():
            k = k[7:] if k.startswith("module.") else k
            sd[k] = v.cpu() if isinstance(v, torch.Tensor) else v
        torch.save(sd, os.path.join(self.output_dir, f"{self.prefix}_{step}.{self.suffix}"))
        if optimizer is not None:
            dump = {"step": step, "optimizer": optimizer.state_dict()}
            name = "train_state_latest.pt" if rewrite_optimizer else f"train_state_{step}.pt"
            torch.save(dump, os.path.join(self.output_dir, name))


@torch.no_grad()
def validate(model, val_batches):
    """train/train_simple_policy.py:271-300"""
    model.eval()
    tot = dict(total_loss=0.0, pos_loss=0.0, pos_l1_loss=0.0, rot_loss=0.0, open_loss=0.0, open_acc=0.0)
    n_ex = n_b = 0
    for batch in val_batches:
        gt = batch["gt_actions"].cpu()
        pred, loss = model(dict(batch), compute_loss=True)
        pred = pred.cpu()
        tot["open_acc"] += ((torch.sigmoid(pred[..., -1]) > 0.5) == gt[..., -1]).float().sum().item()
        tot["pos_l1_loss"] += torch.nn.functional.l1_loss(pred[..., :3].float(), gt[..., :3]).item()
        for k in ("pos", "rot", "open", "total"):
            tot[k + "_loss"] += loss[k].item()
        n_ex += pred.size(0)
        n_b += 1
    model.train()
    out = {k: v / max(n_b, 1) for k, v in tot.items()}
    out["open_acc"] = tot["open_acc"] / max(n_ex, 1)
    return out


def train(config, make_epoch, steps_per_epoch, val_batches=None, device=None, rank=0, world=1, precision="bf16"):
    """`main(config)` of train_simple_policy.py:53-268.  make_epoch(epoch_id) -> iterable of batches for this rank."""
    T = config.TRAIN
    device = device or torch.device("cuda", 0)
    default_gpu = rank == 0
    torch.manual_seed(config.SEED + (rank if world > 1 else 0))
    acc = max(1, int(T.gradient_accumulation_steps))
    num_train_steps = T.num_train_steps if T.num_train_steps is not None else steps_per_epoch * T.num_epochs
    num_epochs = int(math.ceil(num_train_steps / max(1, steps_per_epoch)))
    num_train_steps *= acc if acc > 1 else 1
    num_epochs *= acc if acc > 1 else 1
    ckpt_dir = os.path.join(config.output_dir, "ckpts")
    if default_gpu:
        os.makedirs(os.path.join(config.output_dir, "logs"), exist_ok=True)
        os.makedirs(ckpt_dir, exist_ok=True)
        with open(os.path.join(config.output_dir, "logs", "training_config.yaml"), "w") as f:
            yaml.safe_dump(_plain(config), f)
        fh = logging.FileHandler(os.path.join(config.output_dir, "logs", "log.txt"))
        LOGGER.addHandler(fh)
        LOGGER.setLevel(logging.INFO)
    saver = ModelSaver(ckpt_dir)

    mcfg = config.MODEL.clone()
    model = MODEL_FACTORY[mcfg.model_class](mcfg)
    LOGGER.info("Model: nweights %d nparams %d" % model.num_parameters)

    # ---- resume (train_simple_policy.py:130-173)
    model_ckpt, opt_ckpt, global_step = config.checkpoint, None, 0
    state_file = os.path.join(ckpt_dir, "train_state_latest.pt")
    if os.path.exists(state_file) and T.resume_training:
        opt_ckpt = torch.load(state_file, map_location="cpu")
        latest = os.path.join(ckpt_dir, "model_step_%d.pt" % opt_ckpt["step"])
        if os.path.exists(latest):
            model_ckpt = latest
        global_step = opt_ckpt["step"]
    restart_epoch = global_step // max(1, steps_per_epoch)
    if model_ckpt is not None:
        ck = torch.load(model_ckpt, map_location="cpu")
        own = model.state_dict()
        model.load_state_dict({k: v for k, v in ck.items() if k in own and v.size() == own[k].size()},
                              strict=config.checkpoint_strict_load)
    model = model.to(device).train()
    eng = model.engine()
    eng.precision = precision
    from . import native
    eng.use_native = native.supported(model)
    step_fn = TrainStep(model, T.learning_rate, tuple(T.betas), T.weight_decay, T.grad_norm, T.warmup_steps, num_train_steps)
    step_fn.lr_fn = lambda gs: max(T.learning_rate * lr_decay_rate(gs, _with(T, num_train_steps=num_train_steps)), 1e-8)
    step_fn.global_step = global_step
    if opt_ckpt is not None:
        step_fn.opt.load_state_dict(opt_ckpt["optimizer"])

    running, t_log, n_log = {}, time.time(), 0
    best_val_step, best_val = None, float("inf")
    epoch_id = restart_epoch
    for epoch_id in range(restart_epoch, num_epochs):
        if global_step >= num_train_steps:
            break
        for step, (batch, upcoming) in enumerate(_lookahead(make_epoch(epoch_id))):
            # `upcoming` is copied to the device and tokenised on a side stream while this step runs (native engine).
            # Accumulation windows run on a micro-step counter that survives epoch boundaries (the reference only zeroes
            # gradients after optimizer.step(), train_simple_policy.py:214-243)
            losses = step_fn.accumulate(batch, first=None, next_batch=upcoming if eng.use_native else None)
            stepped = step_fn.micro_step >= acc
            if stepped:
                step_fn.apply(accumulation=acc)
                global_step = step_fn.global_step
                n_log += 1
            if global_step % T.log_steps == 0 and stepped:
                vals = {k: float(v) for k, v in zip(("pos", "rot", "open", "total"), torch.stack([losses[k] for k in ("pos", "rot", "open", "total")]).tolist())}
                running.update({f"loss_{k}": v for k, v in vals.items()})
                dt = time.time() - t_log
                LOGGER.info(f"==============Epoch {epoch_id} Step {global_step}===============")
                LOGGER.info(", ".join("%s:%.4f" % kv for kv in running.items()) + ", grad_norm:%.4f, steps/s:%.2f" % (step_fn.opt.grad_norm(), n_log / max(dt, 1e-9)))
                t_log, n_log = time.time(), 0
            if default_gpu and global_step % T.save_steps == 0 and stepped:
                saver.save(model, global_step, optimizer=step_fn.opt, rewrite_optimizer=True)
            if val_batches is not None and T.val_steps and global_step % T.val_steps == 0 and stepped:
                vm = validate(model, val_batches)
                LOGGER.info("=================Validation=================")
                LOGGER.info(", ".join("%s: %.4f" % kv for kv in vm.items()))
                if vm["pos_loss"] < best_val:
                    best_val, best_val_step = vm["pos_loss"], global_step
            if global_step >= num_train_steps:
                break
    if default_gpu and global_step % T.save_steps != 0:
        saver.save(model, global_step, optimizer=step_fn.opt, rewrite_optimizer=True)
    if val_batches is not None:
        vm = validate(model, val_batches)
        LOGGER.info(", ".join("%s: %.4f" % kv for kv in vm.items()))
        if vm["pos_loss"] < best_val:
            best_val, best_val_step = vm["pos_loss"], global_step
        LOGGER.info(f"Validation: Best loss: {best_val:.4f} at step {best_val_step}")
    return model, step_fn


def _lookahead(it):
    """(item, next item or None) pairs of an iterable"""
    it = iter(it)
    try:
        cur = next(it)
    except StopIteration:
        return
    for nxt in it:
        yield cur, nxt
        cur = nxt
    yield cur, None


def _plain(node):
    return {k: (_plain(v) if isinstance(v, dict) else v) for k, v in node.items()}


def _with(node, **kw):
    n = ConfigNode(_plain(node))
    for k, v in kw.items():
        n[k] = v
    return n


def build_args(argv=None):
    """train_simple_policy.py:314-343"""
    ap = argparse.ArgumentParser()
    ap.add_argument("--exp-config", type=str, required=True, help="path to config yaml containing info about experiment")
    ap.add_argument("--synthetic", type=int, default=0, metavar="STEPS_PER_EPOCH",
                    help="train on synthetic voxel1cm-shaped batches (no dataset offline): steps per epoch")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("opts", default=None, nargs=argparse.REMAINDER,
                    help="Modify config options from command line (use , to separate values in a list)")
    args = ap.parse_args(argv)
    return args, get_config(args.exp_config, args.opts)


def main(argv=None):
    args, config = build_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s", stream=sys.stdout)
    if os.path.exists(config.output_dir) and os.listdir(config.output_dir):
        LOGGER.warning("Output directory (%s) already exists and is not empty." % config.output_dir)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        from .trainer import nccl_env_defaults
        nccl_env_defaults()
        torch.distributed.init_process_group("nccl", device_id=dev)
    if not args.synthetic:
        raise SystemExit("the LMDB keystep dataset reader is outside this path (SURVEY §8f row 2): pass --synthetic N, or call "
                         "train_loop.train(config, make_epoch, steps_per_epoch) with your own batch iterator")
    from . import synthetic as syn
    bs = config.TRAIN.train_batch_size

    def make_epoch(epoch_id):
        for i in range(args.synthetic):
            yield syn.make_batch(seed=config.SEED + 1000 * rank + 7919 * epoch_id + i, batch_size=bs)

    val = [syn.make_batch(seed=config.SEED + 99991, batch_size=min(bs, 8))]
    train(config, make_epoch, args.synthetic, val_batches=val, device=dev, rank=rank, world=world, precision=args.precision)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
