"""numpy restatement of ``deepsulci.sulci_labeling.method.cutting.cutting``.

Test infrastructure (see oracle/__init__.py) — "parity unpinned": upstream
source is absent; the contract below is anchored on the reference call site
pattern_class.py:229-231 (twin: transfer_learning/transfer_learning.py:470):

    ypred_cut = cutting(yscores, df['vert_notcut'], bck2, threshold)
    ypred_cut = [self.sulci_side_list[y] for y in ypred_cut]

i.e. inputs ``[Nvox, C]`` float scores, ``[Nvox]`` elementary-fold ids,
``[Nvox, 3]`` int coords, int threshold (voxel count, README.md:39); output a
length-Nvox sequence of class indices.

FROZEN RULE (north_star item 4 "cutting threshold and
per-elementary-fold majority vote"):
  1. voxel label  = argmax_c score[v, c]            (ties -> lowest c)
  2. per fold: 56-bin histogram of voxel labels
  3. l1 = most frequent label, l2 = runner-up       (ties -> lowest c)
  4. fold is CUT iff count[l2] > threshold
  5. not cut: every voxel of the fold gets l1 (majority vote)
     cut    : every voxel gets whichever of {l1, l2} has the higher score at
              that voxel (tie -> l1); CUT_SPLIT_RULE below.  [UNVERIFIED-
              UPSTREAM: upstream splits geometrically using ``bck``; ``bck``
              is accepted and ignored here.]
"""
import numpy as np

CUT_SPLIT_RULE = "top2-score"


def cutting_ref(y_scores, y_vert, bck2, threshold):
    y_scores = np.asarray(y_scores)
    y_vert = np.asarray(y_vert).astype(np.int64).reshape(-1)
    n = y_vert.shape[0]
    if n == 0:
        return []
    y_scores = y_scores.reshape(n, -1)
    ncls = y_scores.shape[1]
    y_pred = np.argmax(y_scores, axis=1)
    out = np.empty(n, dtype=np.int64)
    for v in np.unique(y_vert):
        idx = np.nonzero(y_vert == v)[0]
        hist = np.bincount(y_pred[idx], minlength=ncls)
        l1 = int(np.argmax(hist))
        h2 = hist.copy()
        h2[l1] = -1
        l2 = int(np.argmax(h2))
        if ncls > 1 and h2[l2] > threshold:
            s1 = y_scores[idx, l1]
            s2 = y_scores[idx, l2]
            out[idx] = np.where(s1 >= s2, l1, l2)
        else:
            out[idx] = l1
    return out.tolist()
