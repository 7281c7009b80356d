"""Time the Cout = 64 weight-gradient layers of the BASELINE volume (run once with B2_NO_WGRAD_HALO=1, once without)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ops = __import__("2022_pauriau_unetsulc_b200.ops", fromlist=["x"])


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for name, cin, cout, dims in (("dec2.conv2", 64, 64, (96, 112, 96)), ("enc0.conv2", 32, 64, (96, 112, 96)),
                              ("enc1.conv1", 64, 64, (48, 56, 48)), ("dec2.conv1", 192, 64, (96, 112, 96)), ("dec0.conv1", 768, 256, (24, 28, 24)),
                              ("dec0.conv2", 256, 256, (24, 28, 24)), ("dec1.conv1", 384, 128, (48, 56, 48)),
                              ("dec1.conv2", 128, 128, (48, 56, 48)), ("enc1.conv2", 64, 128, (48, 56, 48))):
    D, H, W = dims
    x = ops.ActView(torch.randn(1, D, H, W, cin, device="cuda").bfloat16(), 1, D, H, W, cin)
    dy = ops.ActView(torch.randn(1, D, H, W, cout, device="cuda").bfloat16(), 1, D, H, W, cout)
    t = timeit(lambda: ops.conv3d_wgrad(x, dy, cin, cout))
    fl = 2.0 * 27 * cin * cout * D * H * W
    print("%s halo=%s  %.1f us  %.0f TFLOP/s" % (name, "off" if os.environ.get("B2_NO_WGRAD_HALO") else "on", t * 1e3,
                                                 fl / t / 1e9))
