"""Synthetic skeleton volumes / folds.  Test infrastructure.

Mirrors what reference dataset.py:45-88 produces: input fp32 [1,D,H,W] binary,
labels int64 [D,H,W] with -1 background.
"""
import numpy as np
import torch


def synth_volume(shape=(96, 112, 96), n_classes=56, seed=1234, occupancy=0.03):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(shape, generator=g) < occupancy)
    labels = torch.full(shape, -1, dtype=torch.long)
    lab = torch.randint(0, n_classes, shape, generator=g)
    labels[x] = lab[x]
    return x.to(torch.float32).unsqueeze(0), labels


def synth_points(shape=(40, 48, 40), n_classes=56, seed=1234, occupancy=0.03, names=None):
    """bck2 point list + names, the form dataset.py:47-49 consumes."""
    x, labels = synth_volume(shape, n_classes, seed, occupancy)
    pts = torch.nonzero(x[0]).numpy()
    # make sure the bounding box is exactly `shape`
    corners = np.array([[0, 0, 0], [shape[0] - 1, shape[1] - 1, shape[2] - 1]])
    pts = np.concatenate([pts, corners], 0)
    pts = np.unique(pts, axis=0)
    lab = labels[pts[:, 0], pts[:, 1], pts[:, 2]].numpy().copy()
    rng = np.random.RandomState(seed)
    lab[lab < 0] = rng.randint(0, n_classes, size=int((lab < 0).sum()))
    if names is None:
        names = ['S%02d_left' % i for i in range(n_classes)]
    return pts.tolist(), [names[l] for l in lab]


def synth_folds(coords, cell=(12, 14, 12)):
    """Spatially coherent elementary-fold ids: fold=(z//c)*64+(y//c)*8+(x//c)."""
    c = np.asarray(coords)
    return ((c[:, 2] // cell[2]) * 64 + (c[:, 1] // cell[1]) * 8 + (c[:, 0] // cell[0])).astype(np.int64)


def synth_scores(n, n_classes=56, seed=7, sharp=3.0):
    g = torch.Generator().manual_seed(seed)
    return torch.softmax(sharp * torch.randn(n, n_classes, generator=g), dim=1)
