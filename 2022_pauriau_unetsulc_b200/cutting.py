"""``cutting`` — replaces ``deepsulci.sulci_labeling.method.cutting.cutting`` (reference pattern_class.py:229-231):

    ypred_cut = cutting(yscores, df['vert_notcut'], bck2, threshold)      # -> class index per voxel

Runs the integer pass on the GPU (``b2_fold_vote``): per elementary fold a label histogram, top-1 / top-2, cut iff
count(top-2) > threshold, one label per (sub)fold.  ``cutting_multi`` evaluates several thresholds in one pass
(test_thresholds calls cutting |th_range| = 3 times on the same scores).  No CPU fallback.
"""
import numpy as np
import torch

from . import ops


def _device(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("unetsulc_b200.cutting: needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def cutting_multi(y_scores, y_vert, bck2, thresholds, device=None):
    """Returns an int64 numpy array [len(thresholds), Nvox]."""
    dev = _device(device)
    if isinstance(y_scores, torch.Tensor):
        scores = y_scores.to(device=dev, dtype=torch.float32)
    else:
        scores = torch.as_tensor(np.asarray(y_scores, dtype=np.float32)).to(dev)
    n = scores.shape[0]
    if n == 0:
        return np.zeros((len(thresholds), 0), dtype=np.int64)
    scores = scores.reshape(n, -1).contiguous()
    vert = np.asarray(y_vert).reshape(-1)
    _, inv = np.unique(vert, return_inverse=True)          # arbitrary vertex ids -> dense fold ids
    fold = torch.from_numpy(inv.astype(np.int32)).to(dev)
    out = ops.fold_vote(scores, fold, int(inv.max()) + 1, [int(t) for t in thresholds])
    return out.cpu().numpy().astype(np.int64)


def cutting(y_scores, y_vert, bck2, threshold, device=None):
    return cutting_multi(y_scores, y_vert, bck2, [threshold], device)[0].tolist()
