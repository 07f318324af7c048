"""CPU-only check of the *device* per-thread k-mer code (csrc/kmer_roll.cuh compiled for the
host by tests/host_emul/roll_emul.cu) against the oracle: rolling forward/revcomp words,
canonical choice, murmur3 over register words, tile/lead/tail masking."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle as orc
from sourmash_b200.synth import synth_genome

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "roll_emul.cu")


@pytest.fixture(scope="module")
def emul():
    exe = os.path.join(tempfile.gettempdir(), "smb_roll_emul")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I/usr/local/cuda/include", "-x", "c++", SRC, "-o", exe])

    def run(seq, k, W, lead=0):
        with tempfile.TemporaryDirectory() as td:
            fin, fout = os.path.join(td, "in"), os.path.join(td, "out")
            with open(fin, "wb") as fh:
                fh.write(b"G" * lead + bytes(seq))
            subprocess.check_call([exe, str(k), str(W), str(lead), fin, fout])
            return np.fromfile(fout, dtype=np.uint64)
    return run


@pytest.mark.parametrize("k", [1, 3, 4, 5, 8, 15, 16, 17, 21, 24, 31, 32, 33, 47, 48, 51, 63, 64, 65])
def test_roll_matches_oracle(emul, k):
    g = synth_genome(5000 + k, seed=100 + k, n_every=89)
    g[1000:1500] = np.frombuffer(bytes(g[1000:1500]).lower(), dtype=np.uint8)
    g[2000] = ord("R")
    g[2001] = 0
    g[-1] = ord("n")
    want, err = orc.seq_to_hashes(bytes(g), k, force=True, keep_zeros=True)
    assert err is None
    for W, lead in ((16, 0), (64, 5), (128, 15)):
        got = emul(g, k, W, lead)
        assert np.array_equal(got, want), (k, W, lead)


def test_roll_short_and_exact_lengths(emul):
    for k in (21, 31):
        for n in (k - 1, k, k + 1, 15, 16, 17, 47, 48, 49):
            g = synth_genome(max(n, 1), seed=n)[:n]
            if n < k:
                assert len(emul(g, k, 16, 3)) == 0
                continue
            want, _ = orc.seq_to_hashes(bytes(g), k, force=True, keep_zeros=True)
            assert np.array_equal(emul(g, k, 16, 3), want)
