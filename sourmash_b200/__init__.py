"""sourmash_b200 -- B200-native FracMinHash sketching and sorted-hash intersection.

Drop-in for the two hot paths of sourmash (see DESIGN.md / include/sourmash_b200.h)."""
from . import batch  # noqa: F401
from ._lowlevel import ffi, lib  # noqa: F401

__version__ = "0.1.0"
