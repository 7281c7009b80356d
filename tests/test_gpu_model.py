"""End-to-end GPU parity: unetsulc_b200.UNet3D (hand-written sm_100a kernels, bf16 activations, fp32 accumulate)
against the fp32 oracle restatement (oracle/unet3d_ref.py) on the same seeded inputs and weights.

Stated tolerances (measured on B200, round 1; bf16 storage of 28 intermediate tensors + bf16 weights):
  * vs the fp32 oracle ................ logits rel-L2 <= 3e-2 (measured 2.2e-2; the oracle's own bf16-storage
                                        emulation differs from fp32 by 2.3e-2), max-abs <= 0.25*std, loss rel <= 1e-2,
                                        softmax rows sum to 1, argmax agreement on labelled voxels >= 97 %
  * vs the bf16-storage-emulating oracle: logits rel-L2 <= 1.5e-2 (tensor-core accumulation order flips a few bf16
                                        roundings per layer)
  * gradients vs fp32 oracle .......... cosine >= 0.9 per tensor (ReLU-mask and pool-argmax flips dominate, see
                                        tests/_aligned_oracle.py); with masks aligned: rel-L2 <= 2.5e-2 for every
                                        tensor above the first pooling boundary, <= 0.2 below it
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPE = (24, 32, 40)


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cosine(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def _pair(seed=42, n_classes=56):
    import unetsulc_b200
    from oracle.unet3d_ref import UNet3DRef
    torch.manual_seed(seed)
    ref = UNet3DRef(1, n_classes, final_sigmoid=False, interpolate=True, dropout=0.,
                    conv_layer_order='crg', init_channel_number=64)
    # non-trivial affine parameters so that GN gamma/beta paths are exercised
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.add_(0.2 * torch.randn_like(p))
            if "norm" in n and n.endswith("bias"):
                p.add_(0.1 * torch.randn_like(p))
    ours = unetsulc_b200.UNet3D(1, n_classes, final_sigmoid=False, interpolate=True, dropout=0.,
                                conv_layer_order='crg', init_channel_number=64)
    ours.load_state_dict(ref.state_dict())          # same keys, same shapes
    return ref.cuda(), ours.cuda()


def _data(seed=1234, shape=SHAPE, n_classes=56):
    from oracle.synth import synth_volume
    x, labels = synth_volume(shape, n_classes, seed, occupancy=0.05)
    return x.unsqueeze(0).cuda(), labels.unsqueeze(0).cuda()


def test_forward_eval_softmax_and_train_logits():
    ref, ours = _pair()
    x, labels = _data()
    ref.train(); ours.train()
    with torch.no_grad():
        lr_ = ref(x)
        lo = ours(x)
    assert lo.shape == lr_.shape and lo.dtype == torch.float32
    e = rel_l2(lo, lr_)
    mx = float((lo - lr_).abs().max())
    print("train logits rel-L2 %.3e max-abs %.3e std %.3e" % (e, mx, float(lr_.std())))
    assert e < 3e-2
    assert mx < 0.25 * float(lr_.std())
    ref.emulate_bf16_storage = True
    with torch.no_grad():
        le = ref(x)
    ref.emulate_bf16_storage = False
    print("train logits vs bf16-emulating oracle rel-L2 %.3e" % rel_l2(lo, le))
    assert rel_l2(lo, le) < 1.5e-2
    ref.eval(); ours.eval()
    with torch.no_grad():
        pr = ref(x)
        po = ours(x)
    assert float((po.sum(1) - 1).abs().max()) < 1e-4
    m = labels >= 0
    agree = float((po.argmax(1)[m] == pr.argmax(1)[m]).float().mean())
    print("eval softmax rel-L2 %.3e argmax agreement %.4f" % (rel_l2(po, pr), agree))
    assert rel_l2(po, pr) < 3e-2
    assert agree >= 0.97


def test_dense_autograd_path_matches_oracle_grads():
    """model(x) -> torch CrossEntropyLoss -> backward, exactly the reference's training step (training.py:203-212)."""
    ref, ours = _pair()
    x, labels = _data()
    ref.train(); ours.train()
    crit = nn.CrossEntropyLoss(ignore_index=-1)
    loss_r = crit(ref(x), labels); loss_r.backward()
    loss_o = crit(ours(x), labels); loss_o.backward()
    print("loss ref %.6f ours %.6f" % (float(loss_r), float(loss_o)))
    assert abs(float(loss_o) - float(loss_r)) < 1e-2 * abs(float(loss_r))
    for (n, pr), (_, po) in zip(ref.named_parameters(), ours.named_parameters()):
        assert po.grad is not None, n
        c = cosine(po.grad, pr.grad)
        ratio = float(po.grad.norm() / pr.grad.norm())
        print("grad %-45s cos %.4f norm ratio %.3f rel-L2 %.3e" % (n, c, ratio, rel_l2(po.grad, pr.grad)))
        assert c > 0.9, n
        assert 0.8 < ratio < 1.25, n
    assert rel_l2(ours.final_conv.weight.grad, ref.final_conv.weight.grad) < 4e-2


def test_gradients_with_aligned_relu_masks():
    """Backward composition check: oracle (bf16-storage emulation) with its ReLU masks forced to ours."""
    from unetsulc_b200 import models, ops
    from tests._aligned_oracle import run_aligned
    ref, ours = _pair()
    x, labels = _data()
    ref.train(); ours.train()
    save = models._Saved()
    feat = ours._trunk_forward(x, save)
    head = ours.final_conv
    out = ops.head_ce(feat, labels, head.weight.detach(), head.bias.detach(), compute_grad=True)
    grads = ours._trunk_backward(save, out["dx"], [True] * 42)
    mine_r = [rc["r"].dense().float().permute(0, 4, 1, 2, 3) for rc in save.rec]
    loss_r = run_aligned(ref, x, labels, mine_r)
    assert abs(float(out["loss"][0]) - float(loss_r)) < 2e-3 * abs(float(loss_r))
    ref_grads = [p.grad for n, p in ref.named_parameters() if not n.startswith("final_conv")]
    names = [n for n, p in ref.named_parameters() if not n.startswith("final_conv")]
    assert len(ref_grads) == 42
    for i, (n, g, rg) in enumerate(zip(names, grads, ref_grads)):
        e = rel_l2(g, rg)
        print("aligned grad %-45s rel-L2 %.3e" % (n, e))
        below_pool = n.startswith("encoders.0") or n.startswith("encoders.1") or n.startswith("encoders.2")
        assert e < (0.2 if below_pool else 2.5e-2), n
    assert rel_l2(out["dW"], ref.final_conv.weight.grad) < 2e-2
    assert rel_l2(out["db"], ref.final_conv.bias.grad) < 1e-2


def test_fused_loss_path_matches_dense_path():
    ref, ours = _pair()
    x, labels = _data()
    ours.train()
    crit = nn.CrossEntropyLoss(ignore_index=-1)
    out = ours(x)
    loss_d = crit(out, labels); loss_d.backward()
    gd = [p.grad.clone() for p in ours.parameters()]
    preds_d = out.argmax(1)
    ours.zero_grad()
    loss_f, preds_f = ours.loss_and_preds(x, labels)
    loss_f.backward()
    m = labels >= 0
    assert abs(float(loss_f) - float(loss_d)) < 1e-4 * abs(float(loss_d))
    assert torch.equal(preds_f[m].long(), preds_d[m])
    for (n, p), g0 in zip(ours.named_parameters(), gd):
        e = rel_l2(p.grad, g0)
        print("fused vs dense grad %-45s rel-L2 %.3e" % (n, e))
        assert e < 2e-2, n
    # eval: reference val-phase loss = CE(softmax(z))
    ours.eval()
    with torch.no_grad():
        l2, _ = ours.loss_and_preds(x, labels)
        ref_l2 = crit(ours(x), labels)
    assert abs(float(l2) - float(ref_l2)) < 1e-4 * abs(float(ref_l2))


def test_training_step_is_bitwise_deterministic():
    """No float atomics anywhere: two identical steps give bit-identical loss and gradients."""
    ref, ours = _pair()
    x, labels = _data()
    ours.train()
    res = []
    for _ in range(2):
        ours.zero_grad(set_to_none=True)
        loss, preds = ours.loss_and_preds(x, labels)
        loss.backward()
        torch.cuda.synchronize()
        res.append((loss.detach().clone(), preds.clone(), [p.grad.clone() for p in ours.parameters()]))
    assert torch.equal(res[0][0], res[1][0])
    assert torch.equal(res[0][1], res[1][1])
    for (n, _), a, b in zip(ours.named_parameters(), res[0][2], res[1][2]):
        assert torch.equal(a, b), "non-deterministic gradient: %s (max diff %.3e)" % (n, float((a - b).abs().max()))


def test_transfer_learning_freezing_masks():
    """requires_grad masks by name prefix (transfer_learning/transfer_learning.py:330-335)."""
    ref, ours = _pair()
    x, labels = _data()
    ref.train(); ours.train()
    crit = nn.CrossEntropyLoss(ignore_index=-1)
    for layers in (['final_conv'], ['final_conv', 'decoders.2', 'decoders.1', 'decoders.0']):
        for model in (ref, ours):
            model.zero_grad(set_to_none=True)
            for name, p in model.named_parameters():
                p.requires_grad = any(name.startswith(l) for l in layers)
        crit(ref(x), labels).backward()
        lo, _ = ours.loss_and_preds(x, labels)
        lo.backward()
        for (n, pr), (_, po) in zip(ref.named_parameters(), ours.named_parameters()):
            if pr.requires_grad:
                assert po.grad is not None, n
                assert cosine(po.grad, pr.grad) > 0.9, n
            else:
                assert po.grad is None, n


def test_head_swap_deepcopy_state_dict_roundtrip(tmp_path):
    """final_conv replaced after construction (pattern_class.py:364), deepcopy (transfer_learning.py:159), .mdsm."""
    ref, ours = _pair()
    x, labels = _data()
    torch.manual_seed(0)
    new_head = nn.Conv3d(64, 11, 1)
    ours2 = copy.deepcopy(ours)
    ours2.final_conv = copy.deepcopy(new_head).cuda()
    ref2 = copy.deepcopy(ref)
    ref2.final_conv = copy.deepcopy(new_head).cuda()
    ours2.eval(); ref2.eval()
    with torch.no_grad():
        a, b = ours2(x), ref2(x)
    assert a.shape[1] == 11 and rel_l2(a, b) < 3e-2
    # save like pattern_class.py:303-304, load into the ORACLE (= "loads into the reference unchanged")
    ours2.to(torch.device('cpu'))
    path = str(tmp_path / "m_model.mdsm")
    torch.save(ours2.state_dict(), path)
    ref3 = copy.deepcopy(ref2).cpu()
    ref3.load_state_dict(torch.load(path, map_location='cpu'))
    ours2.to(torch.device('cuda'))
    with torch.no_grad():
        c = ours2(x)
    assert rel_l2(c, a) < 1e-6


def test_cpu_input_is_a_hard_error():
    ref, ours = _pair()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ours(torch.zeros(1, 1, 16, 16, 16))


def test_odd_volume_and_batch2():
    ref, ours = _pair()
    from oracle.synth import synth_volume
    xs, ls = [], []
    for s in (1, 2):
        x, l = synth_volume((17, 26, 21), 56, s, occupancy=0.05)
        xs.append(x); ls.append(l)
    x = torch.stack(xs).cuda(); labels = torch.stack(ls).cuda()
    ref.train(); ours.train()
    with torch.no_grad():
        a, b = ours(x), ref(x)
    print("odd volume batch2 rel-L2 %.3e" % rel_l2(a, b))
    assert rel_l2(a, b) < 3e-2


def test_num_conv_chain_head_matches_sequential_reference():
    """final_conv = nn.Sequential of 1x1x1 convs (reference pattern_class.py:357-363, num_conv > 1): the B200 head
    kernels run the composed affine map, autograd on the small matrices splits dW / db over the links.  Checked against
    the oracle carrying the SAME chain: logits, loss, and the gradient of every link (fused step and autograd
    surface)."""
    from unetsulc_b200.pattern_class import make_head
    ref, ours = _pair(n_classes=56)
    torch.manual_seed(7)
    head = make_head(64, 56, 3).cuda()
    ref.final_conv = head
    ours.final_conv = copy.deepcopy(head)
    assert [tuple(p.shape) for p in ours.head_parameters()] == [tuple(p.shape) for p in ref.final_conv.parameters()]
    x, labels = _data()
    ref.train(); ours.train()
    lr_ = ref(x)
    loss_r = F.cross_entropy(lr_, labels, ignore_index=-1)
    loss_r.backward()
    with torch.no_grad():
        lo = ours(x)
    assert rel_l2(lo, lr_.detach()) < 3e-2
    loss, _, _, grads = ours.forward_backward(x, labels)
    assert len(grads) == 42 + 6 and len(ours.ordered_parameters()) == 48
    assert abs(float(loss[0]) - float(loss_r)) < 1e-2 * abs(float(loss_r))
    ref_head = list(ref.final_conv.parameters())
    for g, p in zip(grads[42:], ref_head):
        assert g.shape == p.grad.shape
        assert cosine(g, p.grad) > 0.97 and 0.8 < float(g.norm() / p.grad.norm()) < 1.25
    # autograd surface: same gradients land in .grad of every link
    lo2, _ = ours.loss_and_preds(x, labels)
    lo2.backward()
    for g, p in zip(grads[42:], ours.head_parameters()):
        assert rel_l2(p.grad, g) < 1e-4
    # one fused optimiser step over all 48 tensors
    from unetsulc_b200.optim import SGD
    opt = SGD(ours.ordered_parameters(), lr=1e-2, momentum=0.9)
    before = [p.detach().clone() for p in ours.head_parameters()]
    opt.step(grads=grads)
    torch.cuda.synchronize()
    for b, p, g in zip(before, ours.head_parameters(), grads[42:]):
        assert torch.allclose(p.detach(), b - 1e-2 * g, rtol=1e-5, atol=1e-7)


def test_deferred_weight_gradient_reduction_matches_inline(monkeypatch):
    """B2_WGRAD_REDUCE=defer (b2_conv3d_wgrad_partial + one b2_wgrad_reduce_multi launch) == the inline reduction."""
    from unetsulc_b200 import models
    ref, ours = _pair()
    x, labels = _data()
    ours.train()
    monkeypatch.setattr(models, "_WGRAD_REDUCE", "inline")
    _, _, _, g_inline = ours.forward_backward(x, labels)
    monkeypatch.setattr(models, "_WGRAD_REDUCE", "defer")
    _, _, _, g_defer = ours.forward_backward(x, labels)
    torch.cuda.synchronize()
    for (n, _), a, b in zip(ours.named_parameters(), g_inline, g_defer):
        assert rel_l2(b, a) < 1e-6, n
