"""CPU: the oracle (oracle/multimae_oracle.py) against fixtures recorded from the live reference
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU tests then compare the CUDA path to it."""
import os

import pytest
import torch

from oracle import multimae_oracle as O


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), map_location="cpu", weights_only=False)


def _cfg_from_fixture(fx):
    c = fx["config"]
    cfg = O.make_config(in_domains=tuple(c["in_domains"]))
    cfg.dim, cfg.depth, cfg.heads = c["dim"], c["depth"], c["heads"]
    cfg.dec_dim, cfg.dec_depth, cfg.dec_heads = c["dec_dim"], c["dec_depth"], c["dec_heads"]
    cfg.posemb_grid = c["image_size"] // 16
    return cfg


@pytest.mark.parametrize("name", ["sampler_small.pt", "sampler_cfg2.pt", "sampler_alpha.pt"])
def test_sampler_bit_exact(golden_dir, name):
    fx = _load(golden_dir, name)
    masks, ids_keep, ids_restore = O.sample_masks(fx["shares"], fx["noises"], fx["noise_all"], fx["num_encoded"])
    assert torch.equal(ids_keep, fx["ids_keep"])
    assert torch.equal(ids_restore, fx["ids_restore"])
    for got, ref in zip(masks, fx["task_masks"]):
        assert torch.equal(got, ref)
        assert got.dtype == torch.int64
    # every row keeps exactly num_encoded tokens (rounding fix-up, multimae/multimae.py:208-212)
    assert all(int((torch.cat(masks, 1) == 0).sum(1)[b]) == fx["num_encoded"] for b in range(ids_keep.shape[0]))


@pytest.mark.parametrize("name", ["tiny3.pt", "interp.pt"])
def test_forward_losses_grads(golden_dir, name):
    fx = _load(golden_dir, name)
    cfg = _cfg_from_fixture(fx)
    p = {k: v.clone() for k, v in fx["state_dict"].items()}
    train = O.trainable(p)
    for v in train.values():
        v.requires_grad_(True)
    losses, preds = O.step_losses(p, fx["inputs"], cfg, fx["task_masks"], fx["ids_keep"], fx["ids_restore"])
    for k, ref in fx["preds"].items():
        torch.testing.assert_close(preds[k], ref, rtol=1e-4, atol=1e-5)
    for k, ref in fx["losses"].items():
        torch.testing.assert_close(losses[k], ref, rtol=1e-5, atol=1e-6)
    sum(losses.values()).backward()
    assert set(fx["grads"]) == {k for k, v in train.items() if v.grad is not None}
    for k, ref in fx["grads"].items():
        torch.testing.assert_close(train[k].grad, ref, rtol=2e-4, atol=2e-6, msg=lambda m, k=k: "%s: %s" % (k, m))
    gn = O.grad_norm([v.grad for v in train.values()])
    torch.testing.assert_close(gn, fx["grad_norm"], rtol=1e-5, atol=0)


def test_state_dict_schema_matches_reference(golden_dir):
    fx = _load(golden_dir, "tiny3.pt")
    cfg = _cfg_from_fixture(fx)
    mine = O.init_params(cfg)
    assert set(mine) == set(fx["state_dict"])
    for k, v in fx["state_dict"].items():
        assert tuple(mine[k].shape) == tuple(v.shape), k
    # the frozen sin-cos tables are deterministic: must match the reference's bit for bit up to fp32 rounding
    for k in mine:
        if k.endswith(".pos_emb"):
            torch.testing.assert_close(mine[k], fx["state_dict"][k], rtol=0, atol=1e-6)


def test_empty_mask_and_unmasked_loss():
    pred = torch.randn(2, 3, 32, 32)
    tgt = torch.randn(2, 3, 32, 32)
    zero = torch.zeros(2, 4, dtype=torch.long)
    assert float(O.masked_mse(pred, tgt, zero)) == 0.0                       # criterion.py:100-101
    half = torch.tensor([[1, 0, 0, 0], [0, 0, 0, 0]])
    v = O.masked_mse(pred, tgt, half)                                       # sample 1 has no masked patch -> nanmean skips it
    ref = ((pred[0] - tgt[0]) ** 2).mean(0)[:16, :16].mean()
    torch.testing.assert_close(v, ref)
    torch.testing.assert_close(O.masked_l1(pred, tgt, None), (pred - tgt).abs().mean())
