"""GPU, >= 2 devices (self-skips otherwise; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu`):
real-NCCL data parallelism over subjects.
  * bucketed, overlapped all-reduce == explicit average of the per-rank gradients; weights stay bit-identical
  * replicas start from rank 0's parameters although every process initialised its own network
  * learning() on a cohort whose size is NOT a multiple of the world size: every rank issues the same collectives
    (zero-weight padding step), no hang, identical results dict and weights on every rank
  * test_thresholds() deals graphs round-robin and gathers the per-graph scores: same result as one rank
"""
import contextlib
import io
import os
import random
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import unetsulc_b200
        from unetsulc_b200 import parallel
        from unetsulc_b200.optim import SGD
        from unetsulc_b200.training import UnetTrainingSulciLabelling
        from oracle.synth import synth_volume, synth_folds
        from tests import harness
        res = {}
        # ---- 1. gradients == explicit average, weights in sync, parameters broadcast from rank 0
        torch.manual_seed(100 + rank)                           # different initialisation on every rank
        model = unetsulc_b200.UNet3D(1, 56).to(dev).train()
        assert not parallel.parameters_in_sync(model)
        red = parallel.BucketedGradReducer(model)
        assert parallel.parameters_in_sync(model)
        opt = SGD(model.ordered_parameters(), lr=1e-2, momentum=0.9)
        x, l = synth_volume((32, 40, 32), 56, 1000 + rank, occupancy=0.05)
        x, l = x.unsqueeze(0).to(dev), l.unsqueeze(0).to(dev)
        model.grad_ready_hook = None
        _, _, _, g_local = model.forward_backward(x, l)
        g_avg = []
        for g in g_local:
            t = g.clone()
            dist.all_reduce(t)
            g_avg.append(t / world)
        model.grad_ready_hook = red._on_layer_ready
        red.begin()
        model.forward_backward(x, l, outs=red.outs())
        views = red.finish()
        torch.cuda.synchronize()
        worst = max(float((a - b).abs().max() / (b.abs().max() + 1e-30)) for a, b in zip(views, g_avg))
        opt.step(grads=views)
        res["grad_worst"] = worst
        res["sync_after_step"] = parallel.parameters_in_sync(model)
        model.grad_ready_hook = None
        del model, red, opt
        # ---- 2. learning() on an uneven cohort (5 train subjects on `world` ranks), graphs on and off
        bck2, names, sslist = harness.synthetic_cohort(n_subjects=7, shape=(14, 16, 12), n_classes=6, seed=2)
        files = sorted(bck2)
        for use_graph in (False, True):
            random.seed(7); np.random.seed(7); torch.manual_seed(7 + rank)
            with contextlib.redirect_stdout(io.StringIO()):
                m = UnetTrainingSulciLabelling(files, 'L', cuda=rank, working_path=tmp,
                                               dict_model={'name': 'dp', 'img_size': [24, 24, 24]},
                                               dict_names=names, dict_bck2=bck2, sulci_side_list=sslist)
                m.use_cuda_graph = use_graph
                m.learning(1e-2, 0.9, 3, files[:5], files[5:], batch_size=1, save_results=True)
            res["learn_%d" % use_graph] = (m.results['epoch_loss_train'], m.results['epoch_loss_val'],
                                           m.results['epoch_acc_train'], m.results['epoch_acc_val'])
            res["learn_sync_%d" % use_graph] = parallel.parameters_in_sync(m.model)
            res["learn_steps_%d" % use_graph] = [t["steps"] for t in m.timings["train"]]
        # ---- 3. test_thresholds sharded over the ranks
        rng = np.random.RandomState(0)
        for g in files:
            pts = np.asarray(bck2[g])
            nb = pts * 2 + 1
            perm = rng.permutation(len(pts))
            m.dict_graph_data[g] = {'nbck': nb.tolist(), 'bck2': pts.tolist(), 'names': names[g],
                                    'vert': list(range(len(pts)))}
            m.dict_graph_data[g + '.notcut'] = {'nbck': nb[perm].tolist(), 'bck2': pts[perm].tolist(),
                                                'names': [names[g][i] for i in perm],
                                                'vert': synth_folds(pts[perm], (8, 8, 8)).tolist()}
        m.results = {'threshold_scores': {}}
        with contextlib.redirect_stdout(io.StringIO()):
            m.test_thresholds(files[:5], [g + '.notcut' for g in files[:5]], [5, 50])
        res["thresholds"] = m.results['threshold_scores']
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_data_parallel_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for r in range(world):
        assert out[r]["grad_worst"] < 1e-5 and out[r]["sync_after_step"]
        for ug in (0, 1):
            assert out[r]["learn_sync_%d" % ug]
            assert out[r]["learn_steps_%d" % ug] == [3, 3, 3]          # ceil(5 / 2) on EVERY rank
    assert out[0]["learn_0"] == out[1]["learn_0"] and out[0]["learn_1"] == out[1]["learn_1"]
    assert out[0]["learn_0"] == out[0]["learn_1"]                      # graph replay == eager under data parallelism
    assert out[0]["thresholds"] == out[1]["thresholds"]
    assert len(out[0]["thresholds"][5][0]) == 5
