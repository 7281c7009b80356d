"""Registers this package under the five names the reference imports (reference pattern_class.py:19-23):

    deepsulci.deeptools.models.UNet3D                   -> models.UNet3D           (B200 kernels)
    deepsulci.sulci_labeling.method.cutting.cutting     -> cutting.cutting         (B200 fold-vote kernel)
    deepsulci.sulci_labeling.analyse.stats.esi_score    -> stats.esi_score
    deepsulci.deeptools.early_stopping.EarlyStopping    -> early_stopping.EarlyStopping
    deepsulci.deeptools.dataset.extract_data            -> dataset.extract_data

so that the reference's UNMODIFIED training.py / pattern_class.py / transfer_learning.py run on top of the B200 path
(``install()``), optionally with inert ``soma`` / ``sigraph`` stand-ins when BrainVISA is absent and the point
lists are supplied through ``dict_bck2`` / ``dict_names`` (dataset.py:47-49 then never touches ``aims``).
"""
import sys
import types


def _module(name):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(_module(parent), child, m)
    return m


def install(unet3d=None, cutting=None, esi_score=None, early_stopping=None, extract_data=None, stub_brainvisa=True):
    """Keyword overrides let the test-suite plug the CPU oracle into the same harness."""
    from . import cutting as _cut, dataset as _ds, early_stopping as _es, models as _models, stats as _stats
    import numpy as np
    if not hasattr(np, "Inf"):          # the reference's divide_lr.py / fine_tunning.py use np.Inf (NumPy < 2)
        np.Inf = np.inf
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    _module("deepsulci.deeptools.models").UNet3D = unet3d or _models.UNet3D
    _module("deepsulci.sulci_labeling.method.cutting").cutting = cutting or _cut.cutting
    _module("deepsulci.sulci_labeling.analyse.stats").esi_score = esi_score or _stats.esi_score
    _module("deepsulci.deeptools.early_stopping").EarlyStopping = early_stopping or _es.EarlyStopping
    _module("deepsulci.deeptools.dataset").extract_data = extract_data or _ds.extract_data
    if stub_brainvisa:
        if "soma" not in sys.modules:
            soma = _module("soma")
            aims = _module("soma.aims")

            def _no_aims(*a, **k):
                raise RuntimeError("soma.aims is a stand-in: supply dict_bck2 / dict_names (no BrainVISA here)")
            aims.read = _no_aims
            soma.aims = aims
        if "sigraph" not in sys.modules:
            sg = _module("sigraph")

            class FoldLabelsTranslator(object):
                def readLabels(self, *a):
                    raise RuntimeError("sigraph stand-in: no translation file support without BrainVISA")
            sg.FoldLabelsTranslator = FoldLabelsTranslator
