"""How long does the host take to ENQUEUE one training step (Python + ctypes + allocator) vs the GPU time?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unetsulc_b200
from unetsulc_b200.optim import SGD
from oracle.synth import synth_volume
torch.manual_seed(42)
m = unetsulc_b200.UNet3D(1, 56).cuda().train()
opt = SGD(m.ordered_parameters(), lr=1e-2, momentum=0.9)
x, l = synth_volume((96, 112, 96), 56, 1234, 0.03)
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
print("enqueue %.2f ms/step, total %.2f ms/step (GPU-bound if enqueue << total)" % ((t1 - t0) / K * 1e3, (t2 - t0) / K * 1e3))
