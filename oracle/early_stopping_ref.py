"""Restatement of ``deepsulci.deeptools.early_stopping.EarlyStopping``.

Test infrastructure — "parity unpinned" for the upstream file itself, but the
behaviour is pinned by its two self-declared adaptations in the reference
(divide_lr.py:6-7,38-61 and fine_tunning.py:36-57): same counter / best-score
rule.  Call sites: training.py:166 (ctor ``patience=``), :256-257
(``es_stop(epoch_loss, self.model)``, ``.early_stop``).
"""
import math


class EarlyStoppingRef(object):
    def __init__(self, patience=7, verbose=False, save=False, savepath=''):
        self.patience = patience
        self.verbose = verbose
        self.counter = 0
        self.best_score = None
        self.early_stop = False
        self.val_loss_min = math.inf
        self.save = save
        self.savepath = savepath

    def __call__(self, val_loss, model):
        score = -val_loss
        if self.best_score is None:
            self.best_score = score
        elif score < self.best_score:
            self.counter += 1
            print('EarlyStopping counter: %i out of %i' % (self.counter, self.patience))
            if self.counter >= self.patience:
                self.early_stop = True
        else:
            self.best_score = score
            self.counter = 0
