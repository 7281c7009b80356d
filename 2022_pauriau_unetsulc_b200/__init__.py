"""unetsulc-b200: B200-native (sm_100a) hot path behind ``UnetPatternSulciLabelling``.

The directory name starts with a digit (it mirrors the reference repository's name), so import it with
``importlib.import_module("2022_pauriau_unetsulc_b200")`` or through the alias package ``unetsulc_b200``.
"""
from . import _lib  # noqa: F401  (does not load the .so until an op is called)
from .models import UNet3D  # noqa: F401

__all__ = ["UNet3D"]
