"""CPU oracle for the UNet3D hot path behind UnetPatternSulciLabelling.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
directory; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker.

PARITY UNPINNED: the arithmetic of the path (``UNet3D``, ``cutting``,
``esi_score``, ``EarlyStopping``) lives in the third-party BrainVISA package
``deepsulci`` (GitHub brainvisa/deepsulci, un-vendored, no pinned version; the
only version hint is ``brainvisa-share-5.1`` at reference pattern_class.py:152),
which is absent from /root/reference and from this image.  The reference has
no tests, golden vectors or fixtures for the path.  This oracle is therefore a
restatement anchored on the reference's own call sites (cited per function)
and on BASELINE.json's north_star; every ambiguous upstream choice is a named
constant in ``oracle/unet3d_ref.py``.  Host-side semantics (loss bookkeeping,
LR division, file layout) ARE pinned: tests run the reference's unmodified
``training.py`` / ``pattern_class.py`` against this oracle (tests/harness.py).
"""
