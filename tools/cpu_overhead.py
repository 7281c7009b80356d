"""How long does the host take to ENQUEUE one eager training step (Python + ctypes + allocator) vs the GPU time?
usage: python tools/cpu_overhead.py [D H W] [--profile]"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unetsulc_b200
from unetsulc_b200.optim import SGD
from bench import synth_volume

args = [a for a in sys.argv[1:] if not a.startswith("--")]
shape = tuple(int(a) for a in args[:3]) if len(args) >= 3 else (96, 112, 96)
torch.manual_seed(42)
m = unetsulc_b200.UNet3D(1, 56).cuda().train()
opt = SGD(m.ordered_parameters(), lr=1e-2, momentum=0.9)
x, l = synth_volume(shape, 56, 1234, 0.03)
x, l = x.unsqueeze(0).cuda(), l.unsqueeze(0).cuda()


def step():
    loss, _, _, grads = m.forward_backward(x, l)
    opt.step(grads=grads)


for _ in range(5):
    step()
torch.cuda.synchronize()
K = 20
t0 = time.perf_counter()
for _ in range(K):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("%s: enqueue %.2f ms/step, total %.2f ms/step (GPU-bound if enqueue << total)"
      % (shape, (t1 - t0) / K * 1e3, (t2 - t0) / K * 1e3))
if "--profile" in sys.argv:
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(K):
        step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
