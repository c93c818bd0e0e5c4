"""CPU, authoring container only (needs /root/reference): the unchanged reference script's own `get_model(args)` and
`utils.create_model` build THIS package's model through the overlay, and checkpoints round-trip with the reference model."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, types
sys.path.insert(0, %(root)r)
from multimae_b200 import overlay
overlay.install(%(ref)r)
import run_pretraining_multimae as R            # the reference script, unmodified
import multimae_b200.multimae as mine
args = types.SimpleNamespace(model="pretrain_multimae_base", in_domains=["rgb", "depth", "semseg"],
                             out_domains=["rgb", "depth", "semseg"], patch_size=16, decoder_dim=256, decoder_depth=2,
                             decoder_num_heads=8, decoder_use_task_queries=True, decoder_use_xattn=True,
                             extra_norm_pix_loss=True, num_global_tokens=1, drop_path=0.0)
model = R.get_model(args)
assert type(model) is mine.MultiMAE, type(model)
assert R.SpatialOutputAdapter.__module__.startswith("multimae_b200"), R.SpatialOutputAdapter.__module__
assert R.NativeScaler.__module__.startswith("multimae_b200")
assert R.MaskedMSELoss.__module__.startswith("multimae_b200")
# checkpoint compatibility with the real reference model (state_dict schema, both directions)
for k in [k for k in sys.modules if k == "multimae" or k.startswith("multimae.")]:
    del sys.modules[k]
import multimae.multimae as refmm               # now the reference package itself
from multimae.input_adapters import PatchedInputAdapter, SemSegInputAdapter
from multimae.output_adapters import SpatialOutputAdapter
ins = {"rgb": PatchedInputAdapter(3, 1, 16), "depth": PatchedInputAdapter(1, 1, 16), "semseg": SemSegInputAdapter(133, 4, 16, dim_class_emb=64)}
outs = {k: SpatialOutputAdapter(c, s, 16, dim_tokens=256, depth=2, num_heads=8, task=t, context_tasks=["rgb", "depth", "semseg"])
        for k, (c, s, t) in {"rgb": (3, 1, "rgb"), "depth": (1, 1, "depth"), "semseg": (133, 4, "semseg"), "norm_rgb": (3, 1, "rgb")}.items()}
ref = refmm.pretrain_multimae_base(ins, outs, num_global_tokens=1, drop_path_rate=0.0)
sd_ref, sd_mine = ref.state_dict(), model.state_dict()
assert list(sd_ref.keys()) == list(sd_mine.keys())
assert all(sd_ref[k].shape == sd_mine[k].shape for k in sd_ref)
model.load_state_dict(sd_ref, strict=True)
ref.load_state_dict(model.state_dict(), strict=True)
assert sorted(n for n, p in ref.named_parameters() if p.requires_grad) == sorted(n for n, p in model.named_parameters() if p.requires_grad)
assert ref.no_weight_decay() == model.no_weight_decay()
print("DROPIN_OK", sum(p.numel() for p in model.parameters() if p.requires_grad))
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_reference_script_builds_our_model():
    res = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "ref": REF}], capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "DROPIN_OK 97917" in res.stdout.replace(",", ""), res.stdout[-500:]


STEP_SCRIPT = r'''
import ctypes, sys, types
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(root)r + "/tests")
import torch
from multimae_b200 import _lib as L
from multimae_b200 import functional as Fn

# ---- stand-ins for the GPU: a library stub that validates every call's arguments and computes nothing, zero-filled
# "uninitialised" buffers so that the losses the unchanged script reads are finite, no device synchronisation
class Stub:
    calls = []
    def __getattr__(self, name):
        res, argtypes = L.SIGNATURES[name]
        def fn(*args):
            assert len(args) == len(argtypes), name
            for a, t in zip(args, argtypes):
                if not isinstance(a, type(ctypes.byref(ctypes.c_int()))):
                    t.from_param(a)
            Stub.calls.append(name)
            return 4096 if name.endswith("_bytes") else (L.ABI_VERSION if name == "mmae_abi_version" else (b"" if name == "mmae_last_error" else 0))
        return fn
stub = Stub()
L.lib = lambda: stub
L.current_stream = lambda: 0
Fn._require_cuda = lambda t, what: None
_empty = torch.empty
torch.empty = lambda *a, **k: _empty(*a, **k).zero_()
torch.cuda.synchronize = lambda *a, **k: None

from multimae_b200 import overlay
overlay.install(%(ref)r)
import run_pretraining_multimae as R            # the reference script, unmodified
import utils

args = types.SimpleNamespace(model="pretrain_multimae_base", in_domains=["rgb", "depth", "semseg"],
                             out_domains=["rgb", "depth", "semseg"], patch_size=16, decoder_dim=256, decoder_depth=1,
                             decoder_num_heads=8, decoder_use_task_queries=True, decoder_use_xattn=True,
                             extra_norm_pix_loss=True, num_global_tokens=1, drop_path=0.0,
                             opt="adamw", weight_decay=0.05, lr=1e-4, opt_eps=1e-8, opt_betas=[0.9, 0.95], momentum=0.9,
                             balancer_lr_scale=1.0)
model = R.get_model(args)
loss_balancer = R.NoWeightingStrategy()
optimizer = R.create_optimizer(args, {"model": model, "balancer": loss_balancer})
loss_scaler = R.NativeScaler()                   # = multimae_b200.native_scaler.NativeScalerWithGradNormCount via the overlay
tasks_loss_fn = {d: R.DOMAIN_CONF[d]["loss"](patch_size=16, stride=R.DOMAIN_CONF[d]["stride_level"]) for d in args.out_domains}
tasks_loss_fn["norm_rgb"] = R.DOMAIN_CONF["rgb"]["loss"](patch_size=16, stride=1, norm_pix=True)
g = torch.Generator().manual_seed(0)
def batch():
    return ({"rgb": torch.randn(2, 3, 224, 224, generator=g), "depth": torch.rand(2, 1, 224, 224, generator=g) + 0.5,
             "semseg": torch.randint(0, 133, (2, 56, 56), generator=g)}, None)
stats = R.train_one_epoch(model, [batch(), batch()], tasks_loss_fn, loss_balancer, optimizer, torch.device("cpu"), epoch=0,
                          loss_scaler=loss_scaler, max_norm=None, max_skip_norm=None, start_steps=0,
                          lr_schedule_values=[1e-4, 1e-4], wd_schedule_values=[0.05, 0.05], num_encoded_tokens=98,
                          in_domains=args.in_domains, loss_on_unmasked=False, alphas=1.0, sample_tasks_uniformly=False,
                          standardize_depth=True, extra_norm_pix_loss=True, fp32_output_adapters=["semseg"])
# the three half-precision adapters (rgb, depth, norm_rgb) share ONE context projection GEMM and run the *_ctx heads
for name in ("mmae_sample_masks", "mmae_embed_forward", "mmae_block_forward", "mmae_ctxproj_forward", "mmae_dechead_forward_ctx",
             "mmae_dectail_forward", "mmae_masked_loss_forward", "mmae_masked_loss_backward", "mmae_dectail_backward",
             "mmae_dechead_backward_ctx", "mmae_ctxproj_backward", "mmae_block_backward", "mmae_embed_backward"):
    assert name in Stub.calls, name
assert Stub.calls.count("mmae_ctxproj_forward") == 2 and Stub.calls.count("mmae_dechead_forward_ctx") == 2 * 3
assert Stub.calls.count("mmae_ctxproj_backward") == 2 and Stub.calls.count("mmae_dechead_backward_ctx") == 2 * 3
# every head's backward precedes the shared projection's backward of its step
_bw = [c for c in Stub.calls if c in ("mmae_dechead_backward_ctx", "mmae_ctxproj_backward")]
assert _bw == ["mmae_dechead_backward_ctx"] * 3 + ["mmae_ctxproj_backward"] + ["mmae_dechead_backward_ctx"] * 3 + ["mmae_ctxproj_backward"], _bw
assert "mmae_dechead_forward" not in Stub.calls and "mmae_dechead_backward" not in Stub.calls
# fp32_output_adapters=["semseg"]: that adapter's head / block / tail run through the fp32-tier entry points
# the 12 encoder blocks run chained (hand-offs fused), the one-block decoder transformers as single blocks
assert Stub.calls.count("mmae_block_forward_chain") == 2 * 12 == Stub.calls.count("mmae_block_backward_chain")
assert Stub.calls.count("mmae_block_forward") == 2 * (3 * 1) and Stub.calls.count("mmae_block_f32_forward") == 2 * 1
assert Stub.calls.count("mmae_dechead_f32_forward") == 2 and Stub.calls.count("mmae_dectail_f32_backward") == 2
assert Stub.calls.count("mmae_masked_loss_forward") == 2 * 4
assert Stub.calls.count("mmae_grad_unscale_norm") == 2                 # fused unscale + norm over the flat arena, per step
assert all(p.grad is not None and p.grad.data_ptr() == model.grad_arena().view(n).data_ptr()
           for n, p in model.named_parameters() if p.requires_grad)
assert {"[Epoch] loss", "[Epoch] rgb_loss", "[Epoch] depth_loss", "[Epoch] semseg_loss", "[Epoch] norm_rgb_loss",
        "[Epoch] grad_norm", "[Epoch] loss_scale", "[Epoch] lr"} <= set(stats), sorted(stats)
print("STEP_OK", len(Stub.calls))
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_reference_train_one_epoch_drives_our_modules():
    """The reference's own, unmodified `train_one_epoch` (run_pretraining_multimae.py:458-574) - autocast context, depth
    standardisation, `model(**kwargs)`, the four criteria, loss balancer, NativeScaler call, meters - runs two steps over
    this package's modules with a library stub in place of the GPU: every module-level C-ABI entry point is reached with
    well-formed arguments, in the counts the B model implies."""
    res = subprocess.run([sys.executable, "-c", STEP_SCRIPT % {"root": ROOT, "ref": REF}], capture_output=True, text=True,
                         timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-5000:]
    assert "STEP_OK" in res.stdout, res.stdout[-500:]


FULL_SCRIPT = r'''
import ctypes, os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch
from PIL import Image
from multimae_b200 import _lib as L
from multimae_b200 import functional as Fn

class Stub:                                     # validates every C-ABI call, computes nothing (no GPU here)
    calls = []
    def __getattr__(self, name):
        res, argtypes = L.SIGNATURES[name]
        def fn(*args):
            assert len(args) == len(argtypes), name
            for a, t in zip(args, argtypes):
                if not isinstance(a, type(ctypes.byref(ctypes.c_int()))):
                    t.from_param(a)
            Stub.calls.append(name)
            return 4096 if name.endswith("_bytes") else (L.ABI_VERSION if name == "mmae_abi_version" else (b"" if name == "mmae_last_error" else 0))
        return fn
stub = Stub()
L.lib = lambda: stub
L.current_stream = lambda: 0
Fn._require_cuda = lambda t, what: None
_empty = torch.empty
torch.empty = lambda *a, **k: _empty(*a, **k).zero_()
torch.cuda.synchronize = lambda *a, **k: None

# a tiny multi-task image folder: root/<task>/<class>/<name>.png (utils/dataset_folder.py MultiTaskImageFolder)
root, out = %(data)r, %(out)r
rng = np.random.default_rng(0)
for i in range(4):
    for task, arr in (("rgb", rng.integers(0, 255, (64, 64, 3), dtype=np.uint8)),
                      ("depth", rng.integers(1000, 60000, (64, 64), dtype=np.uint16)),
                      ("semseg", rng.integers(0, 133, (64, 64), dtype=np.uint8))):
        os.makedirs(os.path.join(root, task, "scene"), exist_ok=True)
        Image.fromarray(arr).save(os.path.join(root, task, "scene", "%%04d.png" %% i))

from multimae_b200 import overlay
argv = [os.path.join(%(ref)r, "run_pretraining_multimae.py"), "--data_path", root, "--output_dir", out, "--device", "cpu",
        "--batch_size", "2", "--epochs", %(epochs)r, "--warmup_epochs", "0", "--save_ckpt_freq", "1", "--num_workers", "0",
        "--no_pin_mem", "--decoder_depth", "1", "--standardize_depth", "--fp32_output_adapters", "semseg",
        "--no_log_wandb", "--blr", "1e-4"]
try:
    overlay.main(argv)
except SystemExit as e:
    assert not e.code, e.code
assert Stub.calls.count("mmae_embed_forward") == %(steps)d, Stub.calls.count("mmae_embed_forward")
# the script's own NativeScaler call took the one-pass unscale / norm path over the flat gradient arena
assert Stub.calls.count("mmae_grad_unscale_norm") == %(steps)d, Stub.calls.count("mmae_grad_unscale_norm")
print("FULL_OK", sorted(f for f in os.listdir(out) if f.startswith("checkpoint")))
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_unmodified_script_end_to_end_with_checkpoint_resume(tmp_path):
    """`python -m multimae_b200.overlay run_pretraining_multimae.py ...` in process: the reference's argument parser, data
    pipeline (a 4-image multi-task folder written here), get_model, optimizer factory, LR/WD schedules, train_one_epoch,
    log.txt and utils.save_model run unmodified over this package's modules (library stubbed: no GPU here).  The second
    launch auto-resumes from the checkpoint the first one wrote (model, optimizer and scaler state through
    utils/checkpoint.py:119-134) and trains one more epoch."""
    data, out = str(tmp_path / "data"), str(tmp_path / "out")
    os.makedirs(out)
    for epochs, steps, expect in (("1", 2, ["checkpoint-0.pth"]), ("2", 2, ["checkpoint-0.pth", "checkpoint-1.pth"])):
        res = subprocess.run([sys.executable, "-c", FULL_SCRIPT % {"root": ROOT, "ref": REF, "data": data, "out": out,
                                                                   "epochs": epochs, "steps": steps}],
                             capture_output=True, text=True, timeout=900)
        assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-6000:]
        assert "FULL_OK %s" % expect in res.stdout, res.stdout[-1500:]
        if epochs == "2":
            assert "Auto resume checkpoint" in res.stdout or "Resume checkpoint" in res.stdout, res.stdout[-3000:]
    assert os.path.exists(os.path.join(out, "log.txt"))
