#!/usr/bin/env python
"""bench.py — headline benchmark: UNet3D (in 1, out 56, 'crg', 64 filters) full-training steps per second on
synthetic 1x1x96x112x96 skeleton volumes (BASELINE.json configs[1]; configs[3] under torchrun with N ranks).

  python bench.py --gpus 1 --steps K --warmup W            our arm (hand-written sm_100a kernels)
  python bench.py --impl reference --steps K --warmup W     the reference's CPU path (oracle port, host cores)
  torchrun ... bench.py --gpus N ...                        data parallel over subjects, one rank per GPU

A "step" = forward + CrossEntropyLoss(ignore_index=-1) + backward + SGD(lr 1e-2, momentum 0.9) on ONE volume per
rank (weak scaling: global batch = N).  `value` = volumes/s of the whole job with inputs resident in HBM;
`e2e` = the same through the public training-class call with pinned HOST inputs (H2D inside the timed region) and a
D2H read of the loss every step.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPE = (96, 112, 96)
N_CLASSES = 56
FWD_BWD_GFLOP = 5612.086          # full training step, one volume
IGEMM_GFLOP_PER_STEP = 2 * 1862.109  # 13 fprop + 13 dgrad launches of conv3d_igemm_kernel (Cin >= 32 layers)
METRIC = "training volumes/sec"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/): the
    newest profiles/r*_traffic.json.  Returns (bytes_per_launch or None, source)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return None, None
    try:
        d = json.load(open(files[-1]))
        return float(d["dram_bytes_per_launch"]), os.path.relpath(files[-1], ROOT)
    except Exception:
        return None, None


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def synth_dataset(n, rank, device=None, shape=SHAPE):
    from oracle.synth import synth_volume   # synthetic-input generator only (data, not compute)
    xs, ls = [], []
    for i in range(n):
        x, l = synth_volume(shape, N_CLASSES, 1234 + 100 * rank + i, occupancy=0.03)
        xs.append(x.unsqueeze(0))
        ls.append(l.unsqueeze(0))
    return xs, ls


# ------------------------------------------------------------------------------------------------------------ CPU arm
def cpu_train_steps(steps, warmup, sample_shape):
    """The reference's CPU path for this metric: the oracle port (fp32 PyTorch on the host cores; the reference runs
    torch.device('cpu') when cuda=-1, pattern_class.py:109-110) doing the same training step on a bounded sample."""
    import torch
    from oracle.unet3d_ref import UNet3DRef
    from oracle.synth import synth_volume
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    model = UNet3DRef(1, N_CLASSES)
    opt = torch.optim.SGD(model.parameters(), lr=1e-2, momentum=0.9, weight_decay=0)
    crit = torch.nn.CrossEntropyLoss(ignore_index=-1)
    x, l = synth_volume(sample_shape, N_CLASSES, 1234, occupancy=0.03)
    x, l = x.unsqueeze(0), l.unsqueeze(0)
    model.train()

    def step():
        opt.zero_grad()
        loss = crit(model(x), l)
        loss.backward()
        opt.step()
        return float(loss)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    frac = (sample_shape[0] * sample_shape[1] * sample_shape[2]) / float(SHAPE[0] * SHAPE[1] * SHAPE[2])
    return frac / dt, dt, cores, frac


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = (48, 56, 48)
    vps, dt, cores, frac = cpu_train_steps(args.steps, args.warmup, sample)
    desc = ("oracle port (fp32 PyTorch restatement of the reference's UNet3D CPU path), %d host threads; each step "
            "= one SGD training step on a %dx%dx%d crop (%.4f of a 96x112x96 volume), volumes/s = fraction / step time"
            % (cores, sample[0], sample[1], sample[2], frac))
    line = {
        "impl": "reference", "metric": METRIC, "value": vps, "unit": "volumes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "UNet3D(in=1,out=56,'crg',f=64) full training step (SGD lr 1e-2 momentum 0.9), "
                               "synthetic 1x1x96x112x96 skeleton volumes, batch 1 per rank"},
        "cpu_baseline": {"value": vps, "unit": "volumes/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": vps, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import unetsulc_b200
    from unetsulc_b200 import ops, parallel
    from unetsulc_b200.optim import SGD
    from unetsulc_b200.training import UnetTrainingSulciLabelling

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks, peak_src = load_peaks()

    # the public training class; its model/optimiser are what `learning()` would build (training.py:60-74,140)
    sslist = ["S%02d_left" % i for i in range(N_CLASSES)]
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        trainer = UnetTrainingSulciLabelling([], "L", cuda=local, working_path="/tmp/unetsulc_bench",
                                             dict_model={"name": "bench"}, dict_names={}, dict_bck2={},
                                             sulci_side_list=sslist)
        torch.manual_seed(42)
        trainer.load_network()
    model = trainer.model
    model.train()
    opt = SGD(model.ordered_parameters(), lr=1e-2, momentum=0.9, weight_decay=0)
    reducer = parallel.BucketedGradReducer(model) if world > 1 else None

    n_data = 4
    xs_h, ls_h = synth_dataset(n_data, rank)
    xs_h = [x.pin_memory() for x in xs_h]
    ls_h = [l.pin_memory() for l in ls_h]
    xs_d = [x.to(dev) for x in xs_h]
    ls_d = [l.to(dev) for l in ls_h]
    h2d = xs_h[0].numel() * 4 + ls_h[0].numel() * 8

    # the step is replayed from CUDA graphs: one graph on one rank; under data parallelism one graph segment per
    # gradient bucket with the NCCL all-reduces enqueued eagerly in between (NCCL itself is never captured)
    trainer.use_cuda_graph = not args.no_cuda_graph

    def step_resident(i):
        return trainer.train_step_device(xs_d[i % n_data], ls_d[i % n_data], opt, reducer)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, collective=True):
        """collective=False: rank-local timing (no barrier / all-reduce) for work only one rank does"""
        sync = sync_all if collective else torch.cuda.synchronize
        sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for i in range(steps):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = a.elapsed_time(b)
        if world > 1 and collective:
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        sync()
        return ms / 1e3, wall

    # per-kernel CUDA-event profile of the dominant kernel: taken on eager steps (events cannot be recorded inside a
    # replayed graph), same kernels, same shapes, same stream
    graph_flag, trainer.use_cuda_graph = trainer.use_cuda_graph, False
    for i in range(2):
        step_resident(i)
    torch.cuda.synchronize()
    ops.PROFILE = {}
    l0 = ops.LAUNCHES[0]
    prof_steps = min(args.steps, 5)
    prof_secs, _ = timed(step_resident, prof_steps)
    launches_per_step = (ops.LAUNCHES[0] - l0) / prof_steps
    prof, ops.PROFILE = ops.PROFILE, None
    trainer.use_cuda_graph = graph_flag

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    sampler = ClockSampler(local)
    sampler.start()
    secs, wall = timed(step_resident, args.steps)
    launches = int(round(launches_per_step * args.steps))
    clocks = sampler.stop()
    value = world * args.steps / secs

    # roofline of the dominant kernel (conv3d_igemm_kernel: 26 launches / step, fprop + dgrad of the 13 Cin>=32 convs)
    ig = prof.get("conv3d_igemm", [])
    ig_ms = sum(a.elapsed_time(b) for a, b, _ in ig)
    ig_flop = sum(w for _, _, w in ig)
    achieved = ig_flop / (ig_ms * 1e-3) / 1e12 if ig_ms > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    wg = prof.get("conv3d_wgrad", [])
    wg_ms = sum(a.elapsed_time(b) for a, b, _ in wg)
    traffic, traffic_src = load_traffic()
    roofline = {
        "bound": "tensor", "kernel": "conv3d_igemm_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peak_src + " sustained bf16",
        "launches_per_step": len(ig) / max(prof_steps, 1), "avg_launch_ms": ig_ms / max(len(ig), 1),
        "share_of_step": ig_ms * 1e-3 / prof_secs if prof_secs > 0 else None,
        "wgrad_kernel_tflops": (sum(w for _, _, w in wg) / (wg_ms * 1e-3) / 1e12) if wg_ms > 0 else None,
        "wgrad_share_of_step": wg_ms * 1e-3 / prof_secs if prof_secs > 0 else None,
        "profiled": "%d eager steps (%.3f ms/step) with CUDA events around every conv launch; the timed region "
                    "replays the same step as a CUDA graph" % (prof_steps, prof_secs / prof_steps * 1e3),
        "step_tflops_vs_peak": value / world * FWD_BWD_GFLOP * 1e9 / 1e12 / peak,
    }

    # end to end through the public training-class call: pinned host inputs, H2D + D2H inside the timed region
    def step_e2e(i):
        return trainer.train_step(xs_h[i % n_data], ls_h[i % n_data], opt, reducer)

    for i in range(2):
        step_e2e(i)
    e_secs, _ = timed(step_e2e, args.steps)
    e2e = {"value": world * args.steps / e_secs, "unit": "volumes/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 8, "ms_per_step": e_secs / args.steps * 1e3}

    # secondary metric of BASELINE.json: inference ms per hemisphere = eval forward + Softmax scores gathered at the
    # skeleton voxels + the cutting / fold-vote pass for thresholds [50, 100, 150] (pattern_class.py:177-245), device
    # resident, CUDA events, rank 0 only
    inference = None
    if rank == 0:
        from oracle.synth import synth_folds
        from unetsulc_b200 import cutting as cut_mod
        model.eval()
        xi = xs_d[0]
        idx = torch.nonzero(xi.reshape(-1) > 0).reshape(-1)
        coords = torch.nonzero(xi[0, 0] > 0).cpu().numpy()
        vert = synth_folds(coords, (12, 14, 12))
        import numpy as _np
        _, inv = _np.unique(vert, return_inverse=True)
        fold = torch.from_numpy(inv.astype(_np.int32)).to(dev)
        nf = int(inv.max()) + 1

        def infer(_i):
            with torch.no_grad():
                scores, preds = model.scores_at(xi, idx)
                return ops.fold_vote(scores, fold, nf, [50, 100, 150])

        for i in range(3):
            infer(i)
        isecs, _ = timed(infer, 10, collective=False)   # rank 0 only: no barrier
        inference = {"ms_per_hemi": isecs / 10 * 1e3, "voxels": int(idx.numel()), "folds": nf,
                     "what": "eval forward + softmax gather at skeleton voxels + fold vote for 3 thresholds, "
                             "inputs resident in HBM"}
        model.train()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = (48, 56, 48)
        vps, dt, cores, frac = cpu_train_steps(3, 1, sample)
        cpu = {"value": vps, "unit": "volumes/s", "cores": cores, "kind": "port",
               "sample": "oracle port (fp32 PyTorch, %d host threads): 3 SGD training steps on a %dx%dx%d crop "
                         "(%.4f of a volume) after 1 warm-up; %.2f s/step" % (cores, sample[0], sample[1], sample[2],
                                                                              frac, dt)}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "volumes/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "UNet3D(in=1,out=56,'crg',f=64) full training step (SGD lr 1e-2 momentum 0.9), "
                                   "synthetic 1x1x96x112x96 skeleton volumes, batch 1 per rank",
                       "parallelism": "dp%d over subjects" % world,
                       "cuda_graph": bool(trainer.use_cuda_graph),
                       "l2": "no explicit flush: a step streams ~2 GB of activations (>> 126 MB L2) and cycles "
                             "through %d distinct volumes" % n_data},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "wall_s_timed_region": wall, "inference": inference,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    if os.environ.get("B2_DEBUG_DP") == "1":   # debugging aid: dump every thread's Python stack if the run stalls
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ.get("B2_DEBUG_DP_AFTER", "75")), exit=False)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
