"""Diagnostic: error of the split-operand (bf16x3) tcgen05 convolution and of cuDNN fp32 against an fp64 convolution."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unetsulc_b200
from unetsulc_b200 import ops
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
for (cin, cout, dims) in ((64, 64, (24, 32, 40)), (192, 64, (24, 32, 40)), (768, 256, (12, 14, 12))):
    D, H, W = dims
    x = torch.randn(1, cin, D, H, W, device="cuda")
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * (1.0 / (27 * cin) ** 0.5)
    truth = F.conv3d(x.double(), w.double(), padding=1)
    y32 = F.conv3d(x, w, padding=1).double()
    xn = x.permute(0, 2, 3, 4, 1).contiguous().reshape(-1, cin)
    for terms in (3, 4):
        hi = w.bfloat16().float(); lo = (w - hi).bfloat16().float()
        if terms == 3:
            w3 = torch.cat([hi, hi, lo], 1).contiguous()
            xs = ops.exact_split3(xn, cin, 0, cin, D, H, W)
        else:
            w3 = torch.cat([hi, hi, lo, lo], 1).contiguous()
            xs3 = ops.exact_split3(xn, cin, 0, cin, D, H, W)
            b = xs3.buf.reshape(-1, 3 * cin)
            buf = torch.cat([b, b[:, cin:2 * cin]], 1).contiguous().reshape(1, D, H, W, 4 * cin)
            xs = ops.ActView(buf, 1, D, H, W, 4 * cin)
        wf, _ = ops.pack_conv_weights(w3, want_dgrad=False)
        r = ops.exact_conv(xs, wf, terms * cin, cout, relu=False)
        ye = r.reshape(1, D, H, W, cout).permute(0, 4, 1, 2, 3).double()
        print("Cin %4d Cout %4d %s: split-%d conv rel-L2 %.3e max-abs %.3e | cuDNN fp32 rel-L2 %.3e max-abs %.3e (std %.3f)"
              % (cin, cout, dims, terms, float((ye - truth).norm() / truth.norm()), float((ye - truth).abs().max()),
                 float((y32 - truth).norm() / truth.norm()), float((y32 - truth).abs().max()), float(truth.std())))
