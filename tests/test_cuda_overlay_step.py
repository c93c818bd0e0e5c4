"""GPU: the body of the reference's train_one_epoch (run_pretraining_multimae.py:472-541) restated line by line - the
reference checkout does not travel to the GPU box - over the overlay classes with the REAL library:
`torch.cuda.amp.autocast()`, `NativeScalerWithGradNormCount()` with loss scaling ON (scale 65536), the stock
`torch.optim.AdamW` stepping parameters whose gradients alias the flat arena, `optimizer.zero_grad()` between forward and
backward, the DistributedDataParallel stand-in, `fp32_output_adapters=['semseg']`.  What the unchanged script executes on a
GPU, minus its data pipeline and logger."""
import math
import os

import pytest
import torch

from helpers import formula_fill_, rel_l2

pytestmark = pytest.mark.gpu


def _model(golden_dir):
    from test_cuda_parity import _build_model, _load
    fx = _load(golden_dir, "cuda_small.pt")
    model = _build_model(fx["config"])
    formula_fill_(list(model.named_parameters()))
    return model, fx


def _losses():
    from multimae_b200.criterion import MaskedCrossEntropyLoss, MaskedL1Loss, MaskedMSELoss
    return {"rgb": MaskedMSELoss(patch_size=16, stride=1), "depth": MaskedL1Loss(patch_size=16, stride=1),
            "semseg": MaskedCrossEntropyLoss(patch_size=16, stride=4), "norm_rgb": MaskedMSELoss(patch_size=16, stride=1, norm_pix=True)}


def test_train_one_epoch_sequence_with_loss_scaling_and_stock_adamw(golden_dir):
    from multimae_b200 import multimae as mm
    from multimae_b200 import overlay
    from multimae_b200.native_scaler import NativeScalerWithGradNormCount
    from multimae_b200.optim import FlatAdamW
    dev = torch.device("cuda:0")
    steps, lr = 3, 1e-3
    in_domains = ["rgb", "depth", "semseg"]

    # ------------------------------------------------------------------ path A: the script's sequence (:472-541)
    old = mm.AUTO_OWN_GRADIENTS
    mm.AUTO_OWN_GRADIENTS = True                      # what overlay.install() sets
    try:
        model, fx = _model(golden_dir)
        model = model.to(dev).train()
        model = overlay._IdentityDDP(model, device_ids=[0], find_unused_parameters=False)     # :381
        model_without_ddp = model.module
        optimizer = torch.optim.AdamW([{"params": [p for p in model_without_ddp.parameters() if p.requires_grad], "lr_scale": 1.0}],
                                      lr=lr, weight_decay=0.05, betas=(0.9, 0.95), eps=1e-8)   # utils/optim_factory.py:140-174
        loss_scaler = NativeScalerWithGradNormCount()                                           # :391, enabled: fp16-style scaling
        tasks_loss_fn = _losses()
        x_host = fx["inputs"]
        num_encoded = fx["config"]["num_encoded"]
        log_a = []
        for step in range(steps):
            for group in optimizer.param_groups:                                                # :476-480
                group["lr"] = lr * group["lr_scale"]
            tasks_dict = {task: tensor.to(dev, non_blocking=True) for task, tensor in x_host.items()}   # :482-485
            input_dict = {task: tensor for task, tensor in tasks_dict.items() if task in in_domains}
            torch.manual_seed(100 + step)                                                       # same masks in both paths
            with torch.cuda.amp.autocast():                                                     # :500
                preds, masks = model(input_dict, num_encoded_tokens=num_encoded, alphas=1.0, sample_tasks_uniformly=False,
                                     fp32_output_adapters=["semseg"])
                tasks_dict["norm_rgb"] = tasks_dict["rgb"]                                      # :509-511
                masks["norm_rgb"] = masks.get("rgb", None)
                task_losses = {}
                for task in preds:
                    task_losses[task] = tasks_loss_fn[task](preds[task].float(), tasks_dict[task], mask=masks.get(task, None))
                loss = sum(task_losses.values())
            loss_value = sum(task_losses.values()).item()                                       # :525
            assert math.isfinite(loss_value)
            optimizer.zero_grad()                                                               # :533 (after the forward!)
            grad_norm = loss_scaler(loss, optimizer, clip_grad=None, skip_grad=None, parameters=model.parameters(),
                                    create_graph=False)                                         # :536-537
            loss_scale_value = loss_scaler.state_dict()["scale"]                                # :538
            torch.cuda.synchronize()
            assert loss_scale_value == 65536.0
            # the gradients every parameter holds are views of the flat arena, unscaled; the returned norm is their norm
            arena = model_without_ddp.grad_arena()
            lo, hi = arena.flat.data_ptr(), arena.flat.data_ptr() + arena.flat.numel() * 4
            assert all(lo <= p.grad.data_ptr() < hi for p in model_without_ddp.parameters() if p.requires_grad)
            ref_norm = torch.norm(torch.stack([p.grad.norm() for p in model_without_ddp.parameters() if p.requires_grad]))
            assert abs(float(grad_norm) - float(ref_norm)) < 1e-4 * float(ref_norm)
            log_a.append((loss_value, float(grad_norm)))
            if step == 0:
                params_a = {n: p.detach().clone() for n, p in model_without_ddp.named_parameters()}
    finally:
        mm.AUTO_OWN_GRADIENTS = old

    # ------------------------------------------------------------------ path B: no loss scaling, flat fused AdamW
    model_b, _ = _model(golden_dir)
    model_b = model_b.to(dev).train()
    params_0 = {n: p.detach().clone() for n, p in model_b.named_parameters()}
    opt_b = FlatAdamW(model_b, lr=lr, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    scaler_b = NativeScalerWithGradNormCount(enabled=False).attach_arena(model_b.grad_arena())
    tasks_loss_fn = _losses()
    log_b = []
    for step in range(steps):
        tasks_dict = {task: tensor.to(dev) for task, tensor in x_host.items()}
        torch.manual_seed(100 + step)
        preds, masks = model_b({t: tasks_dict[t] for t in in_domains}, num_encoded_tokens=num_encoded, alphas=1.0,
                               fp32_output_adapters=["semseg"])
        tasks_dict["norm_rgb"] = tasks_dict["rgb"]
        masks["norm_rgb"] = masks.get("rgb", None)
        task_losses = {t: tasks_loss_fn[t](preds[t].float(), tasks_dict[t], mask=masks.get(t, None)) for t in preds}
        loss = sum(task_losses.values())
        gn = scaler_b(loss, opt_b, parameters=None)
        log_b.append((float(loss), float(gn)))
        if step == 0:
            params_b = {n: p.detach().clone() for n, p in model_b.named_parameters()}
    torch.cuda.synchronize()
    # Step 0 starts from identical parameters: power-of-two loss scaling is exact through bf16 operands and fp32
    # accumulation (up to the summation order of the split-K reduce-adds), so loss and gradient norm agree, and the stock
    # AdamW applies the same update as the flat fused one.  Later steps compare loosely: Adam's normalised update turns
    # rounding-level differences of near-zero gradients into lr-sized parameter differences.
    (la, ga), (lb, gb) = log_a[0], log_b[0]
    assert abs(la - lb) < 1e-5 * abs(lb) + 1e-6, (la, lb)
    assert abs(ga - gb) < 1e-5 * abs(gb), (ga, gb)
    upd_a = torch.cat([(params_a[n] - params_0[n]).flatten() for n in params_0])
    upd_b = torch.cat([(params_b[n] - params_0[n]).flatten() for n in params_0])
    assert float(upd_b.norm()) > 0 and rel_l2(upd_a, upd_b) < 1e-3, rel_l2(upd_a, upd_b)
    for step, ((la, ga), (lb, gb)) in enumerate(zip(log_a, log_b)):
        assert abs(la - lb) < 2e-2 * abs(lb), (step, la, lb)
        assert abs(ga - gb) < 5e-2 * abs(gb), (step, ga, gb)
    assert log_a[-1][0] != log_a[0][0]               # the parameters moved between steps
    print("overlay sequence: losses %s, grad norms %s, first AdamW update vs the FlatAdamW path %.2e" %
          ([round(l, 4) for l, _ in log_a], [round(g, 3) for _, g in log_a], rel_l2(upd_a, upd_b)))


def test_flat_adamw_state_dict_round_trip(golden_dir):
    """FlatAdamW.state_dict() has torch.optim.AdamW's layout; save -> load into a fresh optimizer -> identical next step
    (the reference's utils.save_model / auto_load_model contract, utils/checkpoint.py:85,124-133), and a skipped step
    (non-finite gradients) does not advance the bias correction."""
    import io
    from multimae_b200.native_scaler import NativeScalerWithGradNormCount
    from multimae_b200.optim import FlatAdamW
    dev = torch.device("cuda:0")

    def one_step(model, opt, scaler, seed, poison=False):
        x = {k: v.to(dev) for k, v in fx["inputs"].items()}
        torch.manual_seed(seed)
        preds, masks = model(x, num_encoded_tokens=fx["config"]["num_encoded"])
        loss = sum(fns[t](preds[t].float(), x["rgb" if t == "norm_rgb" else t], mask=masks.get("rgb" if t == "norm_rgb" else t))
                   for t in preds)
        if poison:
            loss = loss * float("inf")
        return scaler(loss, opt, parameters=None)

    fns = _losses()
    model, fx = _model(golden_dir)
    model = model.to(dev).train()
    opt = FlatAdamW(model, lr=1e-3)
    scaler = NativeScalerWithGradNormCount(enabled=True).attach_arena(model.grad_arena())
    for s in range(2):
        one_step(model, opt, scaler, s)
    assert float(opt._dyn[1]) == 2.0
    before = opt.flat_params.clone()
    one_step(model, opt, scaler, 7, poison=True)        # non-finite gradients: the step is skipped on the device
    torch.cuda.synchronize()
    assert float(opt._dyn[1]) == 2.0 and torch.equal(opt.flat_params, before)
    assert float(scaler.state_dict()["scale"]) == 32768.0
    buf = io.BytesIO()
    torch.save({"optimizer": opt.state_dict(), "model": model.state_dict(), "scaler": scaler.state_dict()}, buf)
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and len(sd["state"]) == len([p for p in model.parameters() if p.requires_grad])
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}

    model2, _ = _model(golden_dir)
    model2 = model2.to(dev).train()
    opt2 = FlatAdamW(model2, lr=1e-3)
    scaler2 = NativeScalerWithGradNormCount(enabled=True).attach_arena(model2.grad_arena())
    buf.seek(0)
    ck = torch.load(buf, map_location="cpu", weights_only=False)
    model2.load_state_dict(ck["model"])
    opt2.load_state_dict(ck["optimizer"])
    scaler2.load_state_dict(ck["scaler"])
    assert float(opt2._dyn[1]) == 2.0 and rel_l2(opt2.exp_avg, opt.exp_avg) == 0.0
    one_step(model, opt, scaler, 11)
    one_step(model2, opt2, scaler2, 11)
    torch.cuda.synchronize()
    assert torch.equal(opt.flat_params, opt2.flat_params)
    assert float(scaler2.state_dict()["scale"]) == float(scaler.state_dict()["scale"])
    # and the layout is torch.optim.AdamW's: a stock optimizer over the same parameters accepts it
    stock = torch.optim.AdamW([p for p in model2.parameters() if p.requires_grad], lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05)
    stock.load_state_dict(ck["optimizer"])
    assert float(next(iter(stock.state.values()))["step"]) == 2.0


def test_lazy_meters_read_one_step_late(golden_dir):
    """MMAE_LAZY_METERS (n3): the `.item()` reads of train_one_epoch (:525-538) and its per-step `torch.cuda.synchronize()`
    stop draining the device - every call site returns the value it read in the previous step; checkpoints still see the
    exact scaler state."""
    import pickle
    from multimae_b200 import lazy_meters
    from multimae_b200.native_scaler import NativeScalerWithGradNormCount
    from multimae_b200.optim import FlatAdamW
    dev = torch.device("cuda:0")
    model, fx = _model(golden_dir)
    model = model.to(dev).train()
    opt = FlatAdamW(model, lr=1e-3)
    scaler = NativeScalerWithGradNormCount(enabled=True).attach_arena(model.grad_arena())
    fns = _losses()
    x = {k: v.to(dev) for k, v in fx["inputs"].items()}
    real_sync = torch.cuda.synchronize
    assert lazy_meters.install(True)
    try:
        assert torch.cuda.synchronize is not real_sync
        read, true = [], []
        for step in range(4):
            torch.manual_seed(step)
            preds, masks = model(x, num_encoded_tokens=fx["config"]["num_encoded"])
            task_losses = {t: fns[t](preds[t].float(), x["rgb" if t == "norm_rgb" else t], mask=masks.get("rgb" if t == "norm_rgb" else t))
                           for t in preds}
            loss = sum(task_losses.values())
            assert isinstance(loss, lazy_meters.DeferredScalar)
            loss_value = sum(task_losses.values()).item()                          # :525
            per_task = {t: l.item() for t, l in task_losses.items()}               # :526
            grad_norm = scaler(loss, opt, parameters=None)
            scale = scaler.state_dict()["scale"]                                   # :538
            torch.cuda.synchronize()                                               # :540 -> step boundary, no device sync
            read.append((loss_value, per_task["semseg"], grad_norm.item(), scale))
            true.append((loss.detach().clone(), task_losses["semseg"].detach().clone(), grad_norm.detach().clone()))
        real_sync()
        true = [tuple(float(t.as_subclass(torch.Tensor)) for t in row) for row in true]
        assert read[0][:3] == pytest.approx(true[0], rel=1e-6)                     # first step: the real values
        for i in range(1, 4):
            assert read[i][:3] == pytest.approx(true[i - 1], rel=1e-6), i          # afterwards: one step late
        assert all(r[3] == 65536.0 for r in read)
        state = pickle.loads(pickle.dumps(scaler.state_dict()))                    # what utils.save_model stores: exact
        assert type(state) is dict and state["scale"] == 65536.0 and state["_growth_tracker"] == 4
    finally:
        lazy_meters.uninstall()
    assert torch.cuda.synchronize is real_sync


def test_device_feed_hands_out_gpu_batches():
    """MMAE_DEVICE_FEED / MMAE_SYNTHETIC_DATA (n4): a DataLoader wrapped in DeviceFeed yields batches that already live on
    the GPU (copied on a side stream one batch ahead), equal to the host batches."""
    from torch.utils.data import DataLoader
    from multimae_b200.data import DeviceFeed, SyntheticMultiTaskDataset
    ds = SyntheticMultiTaskDataset(40, input_size=64, pool=16)
    host = list(DataLoader(ds, batch_size=8, drop_last=True))
    feed = DeviceFeed(DataLoader(ds, batch_size=8, drop_last=True), device="cuda:0")
    assert len(feed) == len(host) == 5
    n = 0
    for (xd, _), (xh, _) in zip(feed, host):
        for k in xh:
            assert xd[k].is_cuda and torch.equal(xd[k].cpu(), xh[k]), k
            y = xd[k].to("cuda:0", non_blocking=True)                              # the script's own .to(device): a no-op
            assert y.data_ptr() == xd[k].data_ptr()
        n += 1
    assert n == 5
