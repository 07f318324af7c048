"""Bootstrap for processes that run the package against the EMULATED library (emul_lib.py): installs a
``sourmash_b200._lowlevel`` whose library path points at the emulated build, before the package is imported.
Test infrastructure only -- imported by tests/host_emul/emulated_checks.py, never by the package."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import emul_lib  # noqa: E402


def install():
    lib_path = emul_lib.build()
    src_path = os.path.join(ROOT, "sourmash_b200", "_lowlevel.py")
    with open(src_path) as fh:
        source = fh.read()
    marker = 'LIB_PATH = os.path.join(_HERE, "libsourmash_b200.so")'
    assert marker in source
    source = source.replace(marker, "LIB_PATH = %r" % lib_path)
    mod = types.ModuleType("sourmash_b200._lowlevel")
    mod.__file__ = src_path
    mod.__package__ = "sourmash_b200"
    sys.modules["sourmash_b200._lowlevel"] = mod
    exec(compile(source, src_path, "exec"), mod.__dict__)
    return lib_path
