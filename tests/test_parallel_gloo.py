"""CPU: the data-parallel plumbing on world_size-2 gloo (bucket layout, overlap protocol, averaging, metrics)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import unetsulc_b200
    from unetsulc_b200 import parallel
    torch.manual_seed(rank)                    # every process initialises its own network ...
    model = unetsulc_b200.UNet3D(1, 56)
    assert not parallel.parameters_in_sync(model)
    red = parallel.BucketedGradReducer(model)  # ... the reducer broadcasts rank 0's parameters
    assert parallel.parameters_in_sync(model)
    assert sum(f.numel() for f in red.flat) == 16321496
    params = model.ordered_parameters()
    for phase_layers in (None, ['final_conv', 'decoders.2', 'decoders.1', 'decoders.0']):
        for n, p in model.named_parameters():
            p.requires_grad = True if phase_layers is None else any(n.startswith(l) for l in phase_layers)
        red.begin()
        # emulate backward: layers report ready from the head (14) down to 0, gradients written in place
        for layer in range(14, -1, -1):
            idx = [i for i in range(44) if parallel.layer_of_param(i) == layer and params[i].requires_grad]
            for i in idx:
                red.outs()[i].fill_(float(rank + 1) * (i + 1))
            if idx:
                model.grad_ready_hook(layer, [red.outs()[i] for i in idx])
        grads = red.finish()
        for i, p in enumerate(params):
            want = (1 + 2) / 2.0 * (i + 1) if p.requires_grad else 0.0
            assert torch.all(grads[i] == want), (i, float(grads[i].flatten()[0]), want)
    counts = torch.full((3, 4), rank + 1, dtype=torch.int64)
    c, loss, n = parallel.allreduce_metrics(counts, 1.5 * (rank + 1), 2)
    assert torch.all(c == 3) and abs(loss - 4.5) < 1e-12 and n == 4
    q.put(rank)
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert sorted(q.get() for _ in range(2)) == [0, 1]
