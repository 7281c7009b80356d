"""Per-kernel timing at the BASELINE shapes (1x1x96x112x96, f=64): CUDA events on the launching stream,
3 warm-ups, L2 flushed between timed launches.  Prints TFLOP/s for the tcgen05 convs and GB/s for the
bandwidth kernels against MEASURED_PEAKS.json.  Usage: python tools/bench_kernels.py [--scale 1]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unetsulc_b200  # noqa: E402
from unetsulc_b200 import ops  # noqa: E402

PEAKS = {"hbm_gbs": 6540.8, "bf16_tflops": 1678.4}
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    PEAKS.update(json.load(open(pk)))

LAYERS = [  # name, Cin, Cout, level
    ("enc0.conv2", 32, 64, 0), ("enc1.conv1", 64, 64, 1), ("enc1.conv2", 64, 128, 1),
    ("enc2.conv1", 128, 128, 2), ("enc2.conv2", 128, 256, 2), ("enc3.conv1", 256, 256, 3),
    ("enc3.conv2", 256, 512, 3), ("dec0.conv1", 768, 256, 2), ("dec0.conv2", 256, 256, 2),
    ("dec1.conv1", 384, 128, 1), ("dec1.conv2", 128, 128, 1), ("dec2.conv1", 192, 64, 0),
    ("dec2.conv2", 64, 64, 0)]
DIMS = [(96, 112, 96), (48, 56, 48), (24, 28, 24), (12, 14, 12)]

_flush = None


def timeit(fn, iters=5, warm=3):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        _flush.zero_()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None
    rows = []
    tot = {"fprop": [0, 0], "dgrad": [0, 0], "wgrad": [0, 0]}
    for name, cin, cout, lvl in LAYERS:
        if only and only not in name:
            continue
        if only == "bw":
            break
        D, H, W = DIMS[lvl]
        x = ops.ActView(torch.randn(1, D, H, W, cin, device="cuda").to(torch.bfloat16), 1, D, H, W, cin)
        dy = ops.ActView(torch.randn(1, D, H, W, cout, device="cuda").to(torch.bfloat16), 1, D, H, W, cout)
        w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.02
        wf, wd = ops.pack_conv_weights(w)
        y = ops.ActView.alloc(1, D, H, W, cout, "cuda")
        dx = ops.ActView.alloc(1, D, H, W, cin, "cuda")
        flop = 2.0 * D * H * W * 27 * cin * cout
        tf = timeit(lambda: ops.conv3d_igemm(x, wf, y, cin, cout, relu=True))
        td = timeit(lambda: ops.conv3d_igemm(dy, wd, dx, cout, cin, relu=False))
        tw = timeit(lambda: ops.conv3d_wgrad(x, dy, cin, cout))
        for k, t in (("fprop", tf), ("dgrad", td), ("wgrad", tw)):
            tot[k][0] += flop
            tot[k][1] += t
        rows.append((name, flop / 1e9, tf * 1e3, flop / tf / 1e12, td * 1e3, flop / td / 1e12, tw * 1e3,
                     flop / tw / 1e12))
        print("%-11s %7.1f GF | fprop %7.3f ms %6.1f TF/s | dgrad %7.3f ms %6.1f TF/s | wgrad %7.3f ms %6.1f TF/s"
              % rows[-1], flush=True)
    for k, (f, t) in tot.items():
        if t > 0:
            print("TOTAL %s: %.1f GF in %.3f ms = %.1f TF/s (%.1f%% of measured burst %.0f)" %
                  (k, f / 1e9, t * 1e3, f / t / 1e12, 100 * f / t / 1e12 / PEAKS["bf16_tflops"], PEAKS["bf16_tflops"]))
    if only and only != "bw":
        return
    # bandwidth kernels at full resolution
    D, H, W = DIMS[0]
    V = D * H * W
    for C in (32, 64):
        r = ops.ActView(torch.randn(1, D, H, W, C, device="cuda").abs().to(torch.bfloat16), 1, D, H, W, C)
        gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
        t = timeit(lambda: ops.relu_gn_stats(r, 32, 1e-5, gamma, beta))
        print("gn_stats  C=%3d: %.3f ms  %.0f GB/s (read %d MB)" % (C, t * 1e3, V * C * 2 / t / 1e9, V * C * 2 >> 20))
        mr, ss = ops.relu_gn_stats(r, 32, 1e-5, gamma, beta)
        y = ops.ActView.alloc(1, D, H, W, C, "cuda")
        t = timeit(lambda: ops.relu_gn_apply(r, ss, y))
        print("gn_apply  C=%3d: %.3f ms  %.0f GB/s" % (C, t * 1e3, 2 * V * C * 2 / t / 1e9))
        pooled = ops.ActView.alloc(1, D // 2, H // 2, W // 2, C, "cuda")
        t = timeit(lambda: ops.relu_gn_apply(r, ss, y, pooled))
        print("gn_apply+pool C=%3d: %.3f ms  %.0f GB/s" % (C, t * 1e3, (2 + 0.125) * V * C * 2 / t / 1e9))
        t = timeit(lambda: ops.relu_gn_bwd(y, r, 32, gamma, mr))
        print("gn_bwd    C=%3d: %.3f ms  %.0f GB/s (2 passes: 5 tensor sweeps)" % (C, t * 1e3, 5 * V * C * 2 / t / 1e9))
    # the same bandwidth kernels at the coarser levels (launch / tail / latency effects)
    for (C, dims) in ((64, DIMS[1]), (128, DIMS[1]), (128, DIMS[2]), (256, DIMS[2]), (512, DIMS[3])):
        d_, h_, w_ = dims
        Vl = d_ * h_ * w_
        r = ops.ActView(torch.randn(1, d_, h_, w_, C, device="cuda").abs().to(torch.bfloat16), 1, d_, h_, w_, C)
        gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
        mr, ss = ops.relu_gn_stats(r, 32, 1e-5, gamma, beta)
        y = ops.ActView.alloc(1, d_, h_, w_, C, "cuda")
        t = timeit(lambda: ops.relu_gn_apply(r, ss, y))
        acc = torch.zeros(4 * C, dtype=torch.int64, device="cuda")
        t2 = timeit(lambda: ops.relu_gn_bwd_from_stats(acc, y, r, 32, gamma, mr))
        print("level %s C=%3d: gn_apply %.1f us %.0f GB/s | gn_bwd finalize+apply %.1f us %.0f GB/s"
              % (dims, C, t * 1e6, 2 * Vl * C * 2 / t / 1e9, t2 * 1e6, 3 * Vl * C * 2 / t2 / 1e9))
    x = torch.zeros(1, 1, D, H, W, device="cuda"); x[torch.rand_like(x) < 0.03] = 1
    w = torch.randn(32, 1, 3, 3, 3, device="cuda")
    y = ops.ActView.alloc(1, D, H, W, 32, "cuda")
    t = timeit(lambda: ops.conv3d_first_fwd(x, w, y))
    print("conv_first fwd: %.3f ms  %.0f GB/s" % (t * 1e3, (V * 4 + V * 64) / t / 1e9))
    t = timeit(lambda: ops.conv3d_first_wgrad(x, y, 32))
    print("conv_first wgrad: %.3f ms  %.0f GB/s" % (t * 1e3, (V * 4 + V * 64) / t / 1e9))
    coarse = ops.ActView(torch.randn(1, 48, 56, 48, 128, device="cuda").to(torch.bfloat16), 1, 48, 56, 48, 128)
    cat = ops.ActView.alloc(1, D, H, W, 192, "cuda")
    t = timeit(lambda: ops.upcat_fwd(coarse, cat.window(64, 128)))
    print("upcat_fwd 128ch: %.3f ms  %.0f GB/s" % (t * 1e3, (V * 128 * 2 * 1.125) / t / 1e9))
    t = timeit(lambda: ops.upcat_bwd(cat.window(64, 128), 48, 56, 48))
    print("upcat_bwd 128ch: %.3f ms  %.0f GB/s" % (t * 1e3, (V * 128 * 2 * 1.125) / t / 1e9))
    rr = ops.ActView(torch.randn(1, 48, 56, 48, 128, device="cuda").abs().to(torch.bfloat16), 1, 48, 56, 48, 128)
    pool = ops.StatPool("cuda")
    t = timeit(lambda: (pool.reset(), ops.upcat_bwd(cat.window(64, 128), 48, 56, 48, stat_r=rr, pool=pool)))
    print("upcat_bwd 128ch + stats: %.3f ms  %.0f GB/s" % (t * 1e3, (V * 128 * 2 * 1.25) / t / 1e9))
    for (cc, dims) in ((256, (24, 28, 24)), (512, (12, 14, 12))):
        fine = ops.ActView.alloc(1, 2 * dims[0], 2 * dims[1], 2 * dims[2], cc + cc // 2, "cuda")
        fine.buf.normal_()
        t = timeit(lambda: ops.upcat_bwd(fine.window(cc // 2, cc), *dims))
        print("upcat_bwd %dch coarse %s: %.3f ms  %.0f GB/s" % (cc, dims, t * 1e3,
              (8 * dims[0] * dims[1] * dims[2] * cc * 2 * 1.125) / t / 1e9))
    dp = ops.ActView.alloc(1, 48, 56, 48, 64, "cuda")
    t = timeit(lambda: ops.maxpool3d_bwd_add(cat.window(0, 64), cat.window(0, 64), dp))
    print("pool_bwd_add 64ch: %.3f ms  %.0f GB/s" % (t * 1e3, (V * 64 * 2 * 3.125) / t / 1e9))
    feat = ops.ActView(torch.randn(1, D, H, W, 64, device="cuda").to(torch.bfloat16), 1, D, H, W, 64)
    labels = torch.full((1, D, H, W), -1, dtype=torch.long, device="cuda")
    m = torch.rand(1, D, H, W, device="cuda") < 0.03
    labels[m] = torch.randint(0, 56, (int(m.sum()),), device="cuda")
    Wh = torch.randn(56, 64, 1, 1, 1, device="cuda") * 0.1; bh = torch.zeros(56, device="cuda")
    t = timeit(lambda: ops.head_ce(feat, labels, Wh, bh, compute_grad=False))
    print("head_ce fwd: %.3f ms" % (t * 1e3))
    t = timeit(lambda: ops.head_ce(feat, labels, Wh, bh, compute_grad=True))
    print("head_ce fwd+bwd: %.3f ms" % (t * 1e3))
    t = timeit(lambda: ops.head_dense_fwd(feat, Wh, bh, softmax=True))
    print("head_dense_fwd: %.3f ms  %.0f GB/s" % (t * 1e3, (V * 64 * 2 + V * 56 * 4) / t / 1e9))


if __name__ == "__main__":
    main()
