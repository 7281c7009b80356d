"""Layer-by-layer comparison of unetsulc_b200.UNet3D against the fp32 oracle (forward r_i and backward d(conv_i))."""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import unetsulc_b200  # noqa
from unetsulc_b200 import ops, models  # noqa
from oracle.unet3d_ref import UNet3DRef  # noqa
from oracle.synth import synth_volume  # noqa


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (24, 32, 40)
    torch.manual_seed(42)
    ref = UNet3DRef(1, 56).cuda()
    ref.emulate_bf16_storage = os.environ.get('EMUL', '1') == '1'
    ours = unetsulc_b200.UNet3D(1, 56).cuda()
    ours.load_state_dict(ref.state_dict())
    x, labels = synth_volume(shape, 56, 1234, occupancy=0.05)
    x, labels = x.unsqueeze(0).cuda(), labels.unsqueeze(0).cuda()
    ref.train(); ours.train()
    conv_out, conv_grad = {}, {}
    names = []
    for n, m in ref.named_modules():
        if isinstance(m, nn.Conv3d) and m.kernel_size == (3, 3, 3):
            names.append(n)
    import torch.nn.functional as F
    orig_conv3d = F.conv3d
    counter = [0]

    save = models._Saved()
    feat = ours._trunk_forward(x, save)
    force_masks = os.environ.get('FORCE_MASKS', '1') == '1'

    def conv3d_spy(inp, w, b=None, *a, **k):
        out = orig_conv3d(inp, w, b, *a, **k)
        if w.shape[-1] == 3:
            n = names[counter[0]]
            if force_masks:   # make relu'(conv) of the oracle equal to ours (isolates composition bugs from mask flips)
                mine = save.rec[counter[0]]["r"].dense().float().permute(0, 4, 1, 2, 3) > 0
                tgt = torch.where(mine, out.detach().clamp_min(1e-4), out.detach().clamp_max(0.0))
                out = out + (tgt - out).detach()
            counter[0] += 1
            conv_out[n] = out.detach()
            out.register_hook(lambda g, n=n: conv_grad.__setitem__(n, g.detach()))
        return out
    F.conv3d = conv3d_spy
    import oracle.unet3d_ref as oref
    oref.F.conv3d = conv3d_spy
    assert ref.emulate_bf16_storage, "spy works on the functional (emulated) path"
    featgrad = {}
    def _fh(mod, inp, out):
        inp[0].register_hook(lambda g: featgrad.__setitem__("g", g.detach()))
        return None
    ref.final_conv.register_forward_hook(_fh)
    loss_r = nn.functional.cross_entropy(ref(x), labels, ignore_index=-1)
    loss_r.backward()

    # ours, instrumented: capture dr of each layer by wrapping ops.relu_gn_bwd
    drs = []
    orig = ops.relu_gn_bwd

    def wrapped(dy, r, G, gamma, mr, want=True):
        out = orig(dy, r, G, gamma, mr, want)
        drs.append((out[0], dy))
        return out
    ops.relu_gn_bwd = wrapped
    head = ours.final_conv
    out = ops.head_ce(feat, labels, head.weight.detach(), head.bias.detach(), compute_grad=True)
    print("loss ours %.6f ref %.6f" % (float(out["loss"][0]), float(loss_r)))
    grads = ours._trunk_backward(save, out["dx"], [True] * 42)
    torch.cuda.synchronize()
    print("forward: relu(conv_i) vs oracle")
    for i, n in enumerate(names):
        r = save.rec[i]["r"].dense().float().permute(0, 4, 1, 2, 3)
        print("  %-32s rel %.3e" % (n, rel(r, torch.relu(conv_out[n]))))
    dxo = out["dx"].dense().float().permute(0, 4, 1, 2, 3)
    print("dfeat rel %.3e" % rel(dxo, featgrad["g"]))
    lab = (labels >= 0).unsqueeze(1)
    g13 = drs[0][0].dense().float().permute(0, 4, 1, 2, 3)
    o13 = conv_grad[names[13]]
    for nm, m in (("labelled", lab), ("unlabelled", ~lab)):
        mm = m.expand_as(g13)
        print("  dr13 %-10s rel %.3e  share of norm^2 %.3e" % (
            nm, rel(g13[mm], o13[mm]), float((o13[mm] ** 2).sum() / (o13 ** 2).sum())))
    idx = torch.nonzero(labels[0] >= 0)[:3]
    r13 = save.rec[13]["r"].dense().float().permute(0, 4, 1, 2, 3)
    for (d, h, w) in idx.tolist():
        print("voxel", d, h, w)
        print("  mine  ", [round(v, 7) for v in g13[0, :10, d, h, w].tolist()])
        print("  oracle", [round(v, 7) for v in o13[0, :10, d, h, w].tolist()])
        print("  dy    ", [round(v, 7) for v in dxo[0, :10, d, h, w].tolist()])
        print("  dy_o  ", [round(v, 7) for v in featgrad["g"][0, :10, d, h, w].tolist()])
        print("  r     ", [round(v, 5) for v in r13[0, :10, d, h, w].tolist()])
        print("  conv_o", [round(v, 5) for v in conv_out[names[13]][0, :10, d, h, w].tolist()])
    diff = (g13 - o13).abs()
    big = diff > 0.2 * o13.abs().clamp_min(1e-12)
    print("fraction of elements with >20%% error among nonzero oracle: %.4f" % float((big & (o13 != 0)).float().sum() / (o13 != 0).float().sum()))
    print("mask disagreement fraction: %.5f" % float(((g13 != 0) != (o13 != 0)).float().mean()))
    print("backward: d(conv_i) vs oracle   |  weight grad")
    params = dict(ref.named_parameters())
    for k, (dr, dy) in enumerate(drs):
        i = 13 - k
        n = names[i]
        g = dr.dense().float().permute(0, 4, 1, 2, 3)
        gw = grads[3 * i]
        print("  %-32s rel %.3e | dW rel %.3e dgamma %.3e dbeta %.3e" % (
            n, rel(g, conv_grad[n]), rel(gw, params[n + ".weight"].grad),
            rel(grads[3 * i + 1], params[n.replace(".conv", ".norm") + ".weight"].grad),
            rel(grads[3 * i + 2], params[n.replace(".conv", ".norm") + ".bias"].grad)))


if __name__ == "__main__":
    main()
