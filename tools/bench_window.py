"""Does the interleaved concat layout cost convolution throughput?  dec2.conv1-like fprop (K chunks of 64 channels, N = 64,
96x112x96) reading its input as (a) channel windows of one [V][192] buffer (what the zero-copy concat does) and
(b) three dense [V][64] tensors... approximated by one dense [V][64] tensor read three times (same bytes, dense rows)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unetsulc_b200
from unetsulc_b200 import ops
from tools.bench_kernels import timeit
D, H, W = 96, 112, 96
cat = ops.ActView(torch.randn(1, D, H, W, 192, device="cuda").to(torch.bfloat16), 1, D, H, W, 192)
dense = ops.ActView(torch.randn(1, D, H, W, 64, device="cuda").to(torch.bfloat16), 1, D, H, W, 64)
w = torch.randn(64, 64, 3, 3, 3, device="cuda") * 0.02
wf, wd = ops.pack_conv_weights(w)
y = ops.ActView.alloc(1, D, H, W, 64, "cuda")
flop = 2.0 * D * H * W * 27 * 64 * 64
for name, x in (("dense [V][64]", dense), ("window [0:64) of [V][192]", cat.window(0, 64)),
                ("window [64:128) of [V][192]", cat.window(64, 64))):
    t = timeit(lambda: ops.conv3d_igemm(x, wf, y, 64, 64, relu=True))
    print("slab fprop 64->64 %-28s %.3f ms %.0f TF/s" % (name, t * 1e3, flop / t / 1e12))
dy = ops.ActView(torch.randn(1, D, H, W, 64, device="cuda").to(torch.bfloat16), 1, D, H, W, 64)
for name, x in (("dense [V][64]", dense), ("window [0:64) of [V][192]", cat.window(0, 64))):
    t = timeit(lambda: ops.conv3d_wgrad(x, dy, 64, 64))
    print("wgrad 64->64 %-28s %.3f ms %.0f TF/s" % (name, t * 1e3, flop / t / 1e12))
os.environ["B2_NO_SLAB"] = "1"
