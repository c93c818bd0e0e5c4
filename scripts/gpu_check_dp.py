"""2+ GPUs (torchrun): the bucketed in-place NCCL gradient reduction against a single-process recomputation.

Every rank runs one data-parallel step (own batch, own mask draw) through FlatGradReducer; rank 0 then recomputes every
rank's gradients locally (same seeds, reducer detached) and checks  reduced_sum / world == mean of the local gradients,
and that the parameters after the AdamW step are bit-identical on all ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/gpu_check_dp.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from multimae_b200.native_scaler import NativeScalerWithGradNormCount  # noqa: E402
from multimae_b200.optim import FlatAdamW  # noqa: E402
from multimae_b200.parallel import attach_data_parallel, broadcast_parameters  # noqa: E402
from multimae_b200.train_step import TrainStep  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 16


def make(seed_model=0):
    torch.manual_seed(seed_model)
    model, loss_fns = bench.build_model_and_losses(dev)
    return model, loss_fns


def local_step(model, loss_fns, opt, scaler, r):
    """one fwd+bwd of rank r's batch / mask draw on this process; returns the flat gradient"""
    torch.manual_seed(1234 + r)
    x = {k: v.to(dev) for k, v in bench.synthetic_batch(B, 100 * r).items()}
    step = TrainStep(model, loss_fns, opt, scaler, num_encoded_tokens=98, alphas=1.0, loss_sources={"norm_rgb": "rgb"})
    loss, norm = step(x)
    torch.cuda.synchronize()
    return float(loss), float(norm)


# ---- the data-parallel step
model, loss_fns = make()
broadcast_parameters(model)
opt = FlatAdamW(model, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05)
scaler = NativeScalerWithGradNormCount(enabled=False).attach_arena(model.grad_arena())
reducer = attach_data_parallel(model, scaler)
p0 = opt.flat_params.clone()
loss, norm = local_step(model, loss_fns, opt, scaler, rank)
g_dp = model.grad_arena().flat.clone()            # after finish(): SUM over ranks, scaled by 1/world in unscale
p_dp = opt.flat_params.clone()
print("rank %d: loss %.5f grad-norm %.5f buckets %d" % (rank, loss, norm, len(reducer.buckets)), flush=True)

# parameters identical on every rank after the update
ref = p_dp.clone()
dist.broadcast(ref, 0)
same = torch.equal(ref, p_dp)
flags = torch.tensor([1.0 if same else 0.0], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)

ok = True
if rank == 0:
    # ---- recompute every rank's gradient locally, no reducer
    gsum = None
    for r in range(world):
        m2, l2 = make()
        o2 = FlatAdamW(m2, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05)
        s2 = NativeScalerWithGradNormCount(enabled=False).attach_arena(m2.grad_arena())
        assert torch.equal(o2.flat_params, p0), "initial parameters differ from the broadcast ones"
        local_step(m2, l2, o2, s2, r)
        g = m2.grad_arena().flat.clone()
        gsum = g if gsum is None else gsum + g
        o2.release_mirror()
        del m2, o2
    mean = gsum / world
    err = float((g_dp - mean).norm() / mean.norm())
    print("reduced gradient vs mean of locally recomputed gradients: rel L2 %.3e (grad norm %.4f)" % (err, float(mean.norm())))
    ok = err < 2e-3 and bool(flags.item() == 1.0)      # fp32 atomics / split-K order only
    print("parameters identical on all ranks after the step:", bool(flags.item() == 1.0))
    print("DP CHECK", "OK" if ok else "FAILED", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
