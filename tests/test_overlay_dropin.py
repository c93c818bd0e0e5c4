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
