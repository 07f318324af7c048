"""`sourmash sketch dna` driver on the batched GPU path.

Counterpart of the record loop of /root/reference/src/sourmash/command_sketch.py:662-789
(`_compute_individual`: screed record loop -> ``sig.add_sequence(seq, force=not check_sequence)``)
: all records of all input files go to the GPU in one ``smb_sketch_sequences`` call, and one
``SourmashSignature`` per file (or per record with ``singleton=True``) comes back carrying one
sketch per ksize.  The FASTA/FASTQ reader is a minimal host-side parser ("next" row f2 of the
scope table: not accelerated).
"""
import gzip
import os

import numpy as np

from . import batch as B
from ._lowlevel import lib
from .minhash import MinHash
from .signature import SourmashSignature
from ._ffi import rustcall


def read_sequences(path):
    """[(name, sequence bytes)] of a FASTA or FASTQ file (optionally gzipped)."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as fh:
        data = fh.read()
    if not data:
        return []
    if data[:1] == b"@":                                   # FASTQ: 4-line records
        lines = data.split(b"\n")
        return [(lines[i][1:].decode("utf-8", "replace"), lines[i + 1].strip())
                for i in range(0, len(lines) - 1, 4) if lines[i].startswith(b"@")]
    out = []
    for chunk in data.split(b">")[1:]:
        head, _, body = chunk.partition(b"\n")
        out.append((head.strip().decode("utf-8", "replace"), body.replace(b"\n", b"").replace(b"\r", b"")))
    return out


def _signature_from_rows(rows, abunds, ksizes, scaled, num, seed, track, name, filename):
    sig = SourmashSignature.__new__(SourmashSignature)
    sig._objptr = lib.signature_new()
    sig._shared = False
    for i, k in enumerate(ksizes):
        mh = MinHash(num, k, scaled=scaled, seed=seed, track_abundance=track)
        if track:
            mh.set_abundances(dict(zip(rows[i].tolist(), abunds[i].tolist())))
        else:
            mh.add_many(rows[i])
        rustcall(lib.signature_push_mh, sig._get_objptr(), mh._get_objptr())
    if name:
        sig._name = name
    if filename:
        sig.filename = filename
    return sig


def sketch_fasta_files(filenames, *, ksizes=(21, 31, 51), scaled=1000, num=0, seed=42, track_abundance=False,
                       singleton=False, name_from_first=False, check_sequence=False):
    """Sketch every input file on the GPU; returns a list of SourmashSignature (one sketch per ksize).

    ``check_sequence=True`` reproduces ``--check-sequence`` (force=False: the first invalid k-mer
    raises ValueError) through the per-record ABI call; the default skips invalid k-mers like
    the reference CLI (command_sketch.py:827-832)."""
    ksizes = list(ksizes)
    records, owner, names, files = [], [], [], []
    for path in filenames:
        recs = read_sequences(path)
        if singleton:
            for name, seq in recs:
                owner.append(len(names)); names.append(name); files.append(str(path)); records.append(seq)
        else:
            first = recs[0][0] if recs and name_from_first else ""
            for _, seq in recs:
                owner.append(len(names)); records.append(seq)
            names.append(first); files.append(str(path))
    n_sk = len(names)
    if check_sequence:
        sigs = []
        for s in range(n_sk):
            sig = _signature_from_rows([np.zeros(0, np.uint64)] * len(ksizes), [np.zeros(0, np.uint64)] * len(ksizes),
                                       ksizes, scaled, num, seed, track_abundance, names[s], files[s])
            for seq, o in zip(records, owner):
                if o == s:
                    sig.add_sequence(seq, force=False)
            sigs.append(sig)
        return sigs
    lens = np.array([len(r) for r in records], dtype=np.uint64)
    offs = np.zeros(len(records) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    seqs = np.frombuffer(b"".join(records), dtype=np.uint8) if records else np.zeros(0, np.uint8)
    sset, _ = B.sketch_sequences(seqs, offs, ksizes, scaled=scaled, num=num, seed=seed,
                                 track_abundance=track_abundance,
                                 seq_to_sketch=np.array(owner, dtype=np.uint32), n_sketches=n_sk)
    h, off, ab = sset.to_host(with_abunds=True)
    nk = len(ksizes)
    sigs = []
    for s in range(n_sk):
        rows = [h[int(off[s * nk + j]):int(off[s * nk + j + 1])] for j in range(nk)]
        abr = [ab[int(off[s * nk + j]):int(off[s * nk + j + 1])] for j in range(nk)] if ab is not None else None
        sigs.append(_signature_from_rows(rows, abr, ksizes, scaled, num, seed, track_abundance, names[s], files[s]))
    return sigs
