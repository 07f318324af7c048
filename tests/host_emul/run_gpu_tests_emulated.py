"""Runs GPU-marked tests of the suite against the EMULATED library (emul_lib.py) -- a way to exercise the GPU
parity tests when no GPU is at hand (slow: one CPU thread plays the device; pick tests with -k / file names).
Test infrastructure only.

    python tests/host_emul/run_gpu_tests_emulated.py tests/test_gpu_zip.py -x -q
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emulated_boot  # noqa: E402

emulated_boot.install()

import pytest  # noqa: E402

sys.exit(pytest.main(["-m", "gpu", "-p", "no:cacheprovider"] + sys.argv[1:]))
