"""numpy restatement of ``deepsulci.sulci_labeling.analyse.stats.esi_score``.

Test infrastructure — "parity unpinned".  Signature anchored on reference
training.py:223-225 (int labels) and pattern_class.py:233-234 (name labels):
``esi_score(y_true, y_pred, labels)``; accuracy reported = 1 - ESI.

ESI = sum_l (FP_l + FN_l) / sum_l (FP_l + FN_l + 2 TP_l), l over ``labels``
(formula of the paper cited at README.md:3) [UNVERIFIED-UPSTREAM].
"""
import numpy as np


def esi_counts_ref(y_true, y_pred, labels):
    y_true = np.asarray(y_true)
    y_pred = np.asarray(y_pred)
    tp, fp, fn = [], [], []
    for l in labels:
        t = (y_true == l)
        p = (y_pred == l)
        tp.append(int(np.sum(t & p)))
        fp.append(int(np.sum(~t & p)))
        fn.append(int(np.sum(t & ~p)))
    return np.array(tp, np.int64), np.array(fp, np.int64), np.array(fn, np.int64)


def esi_score_ref(y_true, y_pred, labels):
    tp, fp, fn = esi_counts_ref(y_true, y_pred, labels)
    num = float(np.sum(fp + fn))
    den = float(np.sum(fp + fn + 2 * tp))
    if den == 0:
        return 0.0
    return num / den
