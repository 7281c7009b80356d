"""``esi_score`` — replaces ``deepsulci.sulci_labeling.analyse.stats.esi_score`` (reference training.py:223-225,
pattern_class.py:233-234).  ESI = sum_l (FP_l + FN_l) / sum_l (FP_l + FN_l + 2 TP_l) over ``labels``.

Device path: int32 label vectors on the GPU are counted by ``b2_esi_counts`` (integer atomics, exact); the host
signature with Python lists / names (what the reference passes) is counted with numpy — it is host bookkeeping in
the reference too.
"""
import numpy as np
import torch

from . import ops


def esi_from_counts(counts, labels):
    """counts: [3, C] (TP, FP, FN) array-like; labels: iterable of class indices to include."""
    c = counts.detach().cpu().numpy() if isinstance(counts, torch.Tensor) else np.asarray(counts)
    idx = np.asarray([l for l in labels if 0 <= l < c.shape[1]], dtype=np.int64)
    tp, fp, fn = (c[k][idx].astype(np.float64).sum() for k in range(3))
    den = fp + fn + 2.0 * tp
    return float((fp + fn) / den) if den > 0 else 0.0


def esi_score(y_true, y_pred, labels):
    if isinstance(y_true, torch.Tensor) and y_true.is_cuda:
        n_classes = int(max(labels)) + 1 if len(labels) else 1
        counts = ops.esi_counts(y_true.to(torch.int32), y_pred.to(torch.int32), n_classes)
        return esi_from_counts(counts, labels)
    yt = np.asarray(y_true)
    yp = np.asarray(y_pred)
    num = den = 0.0
    for l in labels:
        t = (yt == l)
        p = (yp == l)
        tp = float(np.sum(t & p))
        fp = float(np.sum(~t & p))
        fn = float(np.sum(t & ~p))
        num += fp + fn
        den += fp + fn + 2.0 * tp
    return num / den if den > 0 else 0.0
