"""Multi-GPU check (run under torchrun on >= 2 GPUs): data-parallel step == average of the per-rank gradients,
weights stay in sync.  Usage: torchrun --nproc-per-node 2 tests/gpu_dp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unetsulc_b200  # noqa: E402
from unetsulc_b200 import parallel  # noqa: E402
from unetsulc_b200.optim import SGD  # noqa: E402
from oracle.synth import synth_volume  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(42)
    model = unetsulc_b200.UNet3D(1, 56).to(dev).train()
    red = parallel.BucketedGradReducer(model)
    opt = SGD(model.ordered_parameters(), lr=1e-2, momentum=0.9)
    x, l = synth_volume((32, 40, 32), 56, 1000 + rank, occupancy=0.05)
    x, l = x.unsqueeze(0).to(dev), l.unsqueeze(0).to(dev)
    # reference: local gradients without reduction, averaged explicitly
    model.grad_ready_hook = None
    _, _, _, g_local = model.forward_backward(x, l)
    g_avg = []
    for g in g_local:
        t = g.clone()
        dist.all_reduce(t)
        g_avg.append(t / world)
    # data-parallel path: bucketed, overlapped all-reduce (AVG) into the flat buckets
    model.grad_ready_hook = red._on_layer_ready
    red.begin()
    _, _, _, grads = model.forward_backward(x, l, outs=red.outs())
    views = red.finish()
    torch.cuda.synchronize()
    worst = 0.0
    for a, b in zip(views, g_avg):
        worst = max(worst, float((a - b).abs().max() / (b.abs().max() + 1e-30)))
    opt.step(grads=views)
    # weights identical on all ranks after the step
    chk = torch.stack([p.detach().double().sum() for p in model.parameters()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    insync = bool(torch.equal(lo, hi))
    if rank == 0:
        print("DP check: world %d, max rel diff vs explicit average %.3e, weights in sync: %s" % (world, worst, insync))
    ok = worst < 1e-5 and insync
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
