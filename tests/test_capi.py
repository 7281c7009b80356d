"""CPU: the C-ABI shared library loads and exports exactly the symbols include/unetsulc_b200.h declares."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "unetsulc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", text)))


def test_header_and_ctypes_table_agree():
    import unetsulc_b200
    from unetsulc_b200 import _lib
    assert _declared() == sorted(_lib.SIGNATURES)


def test_library_loads_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from unetsulc_b200 import _lib
    lib = _lib.load()
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (b2_[a-z0-9_]+)", out))
    for name in _declared():
        assert name in exported, name
        assert getattr(lib, name) is not None
    assert lib.b2_last_error() is not None
    # size queries are host-only arithmetic: callable without a GPU
    assert lib.b2_gn_workspace_bytes(1, 64) == 296 * 64 * 2 * 4
    assert lib.b2_relu_gn_bwd_workspace_bytes(2, 64) == 2 * 296 * 64 * 2 * 4 + 2 * 64 * 6 * 4
    assert lib.b2_conv3d_first_wgrad_workspace_bytes(32) == 592 * 27 * 32 * 4
    assert lib.b2_head_workspace_bytes(64) > 0
    assert lib.b2_fold_vote_workspace_bytes(100, 56, 7, 3) > 0


def test_sass_contains_tcgen05_and_tma():
    from unetsulc_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass          # tcgen05.mma
    assert "UTMALDG" in sass          # cp.async.bulk.tensor
    assert "LDTM" in sass             # tcgen05.ld
    assert "HMMA." not in sass.replace("UTCHMMA", "")   # no legacy mma.sync path


def test_wgrad_plan_of_the_baseline_layers():
    """The weight-gradient launch plan, read back on the CPU through the workspace size (= split slices x 27 x Cin x
    Cout fp32; 148 SMs assumed without a device): halo mode cuts Cout <= 128 layers into 64-channel N tiles with 9
    shifted slots per chunk, decoders.0.conv1 (81 columns > 74) takes stream-K runs with three partial slices."""
    from unetsulc_b200 import _lib
    lib = _lib.load()
    expect = {
        # (Cin, Cout, D, H, W): split slices
        (64, 64, 96, 112, 96): 49,      # halo: 5 groups as 2 + 2 + 1 chunks -> 3 columns
        (32, 64, 96, 112, 96): 74,      # halo, 32-channel slots: 3 groups as 2 + 1
        (192, 64, 96, 112, 96): 21,     # roles swapped (dY shifted), N = 192: 7 columns
        (384, 128, 48, 56, 48): 5,      # halo, two 64-channel N tiles x 14 chunks
        (128, 128, 48, 56, 48): 14,
        (64, 128, 48, 56, 48): 24,
        (768, 256, 24, 28, 24): 3,      # stream-K
        (256, 256, 24, 28, 24): 5,
        (256, 512, 12, 14, 12): 2,
    }
    for (cin, cout, d, h, w), splits in expect.items():
        nbytes = lib.b2_conv3d_wgrad_workspace_bytes(1, d, h, w, cin, cout)
        assert nbytes == splits * 27 * cin * cout * 4, (cin, cout, nbytes // (27 * cin * cout * 4))
    assert lib.b2_conv3d_wgrad_workspace_bytes(1, 8, 8, 8, 48, 64) == -1   # Cin must be a multiple of 32
