"""Plateau trackers used by the training loops.

``EarlyStopping`` replaces ``deepsulci.deeptools.early_stopping.EarlyStopping`` (reference training.py:166,256-257);
``DivideLr`` and ``FineTunning`` keep the behaviour of the reference's divide_lr.py:38-61 and fine_tunning.py:36-57
(restated: the originals use ``np.Inf``, which NumPy 2 removed).  All three share one rule: a call "improves" when
``-val_loss >= best``; ``patience`` consecutive non-improving calls fire the tracker's action.
"""
import math
import os.path as op

import torch


class _Plateau(object):
    label = "Plateau"

    def __init__(self, patience=7, verbose=False, save=False, savepath=''):
        self.patience = patience
        self.verbose = verbose
        self.counter = 0
        self.best_score = None
        self.val_loss_min = math.inf
        self.save = save
        self.savepath = savepath

    def _observe(self, val_loss, model):
        """True when the patience is exhausted by this call."""
        score = -val_loss
        if self.best_score is None or not (score < self.best_score):
            self.best_score = score
            self.counter = 0
            if self.save:
                self.save_checkpoint(val_loss, model)
            return False
        self.counter += 1
        print('%s counter: %i out of %i' % (self.label, self.counter, self.patience))
        return self.counter >= self.patience

    def save_checkpoint(self, val_loss, model):
        if self.verbose:
            print('Validation loss decreased (%.6f -> %.6f). Saving model...' % (self.val_loss_min, val_loss))
        torch.save(model.state_dict(), op.join(self.savepath, 'checkpoint.pt'))
        self.val_loss_min = val_loss


class EarlyStopping(_Plateau):
    label = "EarlyStopping"

    def __init__(self, patience=7, verbose=False, save=False, savepath=''):
        super().__init__(patience, verbose, save, savepath)
        self.early_stop = False

    def __call__(self, val_loss, model):
        # the counter is NOT reset when it fires: once stopped it stays stopped until an improvement
        score = -val_loss
        if self.best_score is None:
            self.best_score = score
        elif score < self.best_score:
            self.counter += 1
            print('%s counter: %i out of %i' % (self.label, self.counter, self.patience))
            if self.counter >= self.patience:
                self.early_stop = True
        else:
            self.best_score = score
            self.counter = 0
            if self.save:
                self.save_checkpoint(val_loss, model)


class DivideLr(_Plateau):
    """Flags ``divide_lr`` for one call each time the patience runs out, at most ``repeat`` times."""
    label = "DivideLr"

    def __init__(self, patience=7, verbose=False, save=False, savepath='', repeat=1):
        super().__init__(patience, verbose, save, savepath)
        self.repeat = repeat
        self.stop = False
        self.divide_lr = False

    def __call__(self, val_loss, model):
        self.divide_lr = False
        if self.stop:
            return
        if self._observe(val_loss, model):
            self.divide_lr = True
            self.repeat -= 1
            self.counter = 0
        if self.repeat <= 0:
            self.stop = True


class FineTunning(_Plateau):
    """Fires ``ft_start`` exactly once (then ``stop``)."""
    label = "FineTunning"

    def __init__(self, patience=7, verbose=False, save=False, savepath=''):
        super().__init__(patience, verbose, save, savepath)
        self.stop = False
        self.ft_start = False

    def __call__(self, val_loss, model):
        if self.stop:
            self.ft_start = False
            return
        if self._observe(val_loss, model):
            self.ft_start = True
            self.stop = True
