"""ctypes loader for the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  The product (``sourmash_b200``) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


def _load():
    build()
    lib = C.CDLL(_LIB)
    sz = C.c_size_t
    lib.orc_hash_murmur.restype = C.c_uint64
    lib.orc_hash_murmur.argtypes = [C.c_char_p, sz, C.c_uint64]
    lib.orc_max_hash_for_scaled.restype = C.c_uint64
    lib.orc_max_hash_for_scaled.argtypes = [C.c_uint64]
    lib.orc_scaled_for_max_hash.restype = C.c_uint64
    lib.orc_scaled_for_max_hash.argtypes = [C.c_uint64]
    lib.orc_seq_to_hashes.restype = C.c_int64
    lib.orc_seq_to_hashes.argtypes = [C.c_char_p, sz, C.c_uint32, C.c_uint64, C.c_int, C.c_int,
                                      u64p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.orc_mh_new.restype = C.c_void_p
    lib.orc_mh_new.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_uint32]
    lib.orc_mh_free.argtypes = [C.c_void_p]
    lib.orc_mh_size.restype = sz
    lib.orc_mh_size.argtypes = [C.c_void_p]
    lib.orc_mh_mins.restype = C.POINTER(C.c_uint64)
    lib.orc_mh_mins.argtypes = [C.c_void_p]
    lib.orc_mh_abunds.restype = C.POINTER(C.c_uint64)
    lib.orc_mh_abunds.argtypes = [C.c_void_p]
    lib.orc_mh_max_hash.restype = C.c_uint64
    lib.orc_mh_max_hash.argtypes = [C.c_void_p]
    lib.orc_mh_clear.argtypes = [C.c_void_p]
    lib.orc_mh_remove_hash.argtypes = [C.c_void_p, C.c_uint64]
    lib.orc_mh_add_hash.argtypes = [C.c_void_p, C.c_uint64]
    lib.orc_mh_add_hash_with_abundance.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    lib.orc_mh_add_many.argtypes = [C.c_void_p, u64p, sz]
    lib.orc_mh_add_sequence.restype = C.c_int64
    lib.orc_mh_add_sequence.argtypes = [C.c_void_p, C.c_char_p, sz, C.c_int]
    lib.orc_mh_merge.argtypes = [C.c_void_p, C.c_void_p]
    lib.orc_md5sum.argtypes = [C.c_uint32, u64p, sz, C.c_char_p]
    lib.orc_count_common.restype = C.c_uint64
    lib.orc_count_common.argtypes = [u64p, sz, u64p, sz]
    lib.orc_intersection_size.argtypes = [u64p, sz, u64p, sz, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.orc_intersection_size_num.argtypes = [u64p, sz, u64p, sz, C.c_uint32,
                                              C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.orc_jaccard.restype = C.c_double
    lib.orc_jaccard.argtypes = [u64p, sz, u64p, sz, C.c_uint32]
    lib.orc_angular_similarity.restype = C.c_double
    lib.orc_angular_similarity.argtypes = [u64p, u64p, sz, u64p, u64p, sz]
    lib.orc_downsample_count.restype = sz
    lib.orc_downsample_count.argtypes = [u64p, sz, C.c_uint64]
    lib.orc_compare_all_pairs.argtypes = [u64p, u64p, sz, C.c_uint32, sz, sz, f64p, C.c_int]
    lib.orc_pairwise_common.argtypes = [u64p, u64p, sz, sz, sz, u32p, C.c_int]
    lib.orc_one_vs_many.argtypes = [u64p, sz, u64p, u64p, sz, u64p, C.c_int]
    lib.orc_one_vs_many_bsearch.argtypes = [u64p, sz, u64p, u64p, sz, u64p, C.c_int]
    lib.orc_sketch_scaled.restype = sz
    lib.orc_sketch_scaled.argtypes = [u8p, sz, C.c_uint32, C.c_uint64, C.c_uint64, u64p, sz,
                                      C.POINTER(C.c_uint64)]
    lib.orc_sketch_batch.argtypes = [u8p, u64p, sz, C.c_uint32, C.c_uint64, C.c_uint64, u64p,
                                     u64p, u64p, C.c_int]
    lib.orc_translate_codon.restype = C.c_uint8
    lib.orc_translate_codon.argtypes = [C.c_char_p, sz]
    lib.orc_aa_to_dayhoff.restype = C.c_uint8
    lib.orc_aa_to_dayhoff.argtypes = [C.c_uint8]
    lib.orc_aa_to_hp.restype = C.c_uint8
    lib.orc_aa_to_hp.argtypes = [C.c_uint8]
    for fn in (lib.orc_seq_to_hashes_protein, lib.orc_seq_to_hashes_translate):
        fn.restype = C.c_int64
        fn.argtypes = [C.c_char_p, sz, C.c_uint32, C.c_uint64, C.c_int, C.c_int, u64p]
    lib.orc_mh_add_protein_family.restype = C.c_int64
    lib.orc_mh_add_protein_family.argtypes = [C.c_void_p, C.c_char_p, sz, C.c_int, C.c_int]
    return lib


lib = _load()

HASH_FUNCTIONS = {"dna": 1, "DNA": 1, "protein": 2, "dayhoff": 3, "hp": 4}


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def hash_murmur(data, seed=42):
    if isinstance(data, str):
        data = data.encode()
    return int(lib.orc_hash_murmur(data, len(data), seed))


def max_hash_for_scaled(scaled):
    return int(lib.orc_max_hash_for_scaled(scaled))


def seq_to_hashes(seq, ksize, seed=42, force=False, keep_zeros=False):
    """Returns (hashes ndarray, err_index or None)."""
    if isinstance(seq, str):
        seq = seq.encode()
    n = max(len(seq) - ksize + 1, 1)
    out = np.zeros(n, dtype=np.uint64)
    err = C.c_int64(-1)
    nb = C.c_int64(0)
    r = lib.orc_seq_to_hashes(seq, len(seq), ksize, seed, int(force), int(keep_zeros), out,
                              C.byref(err), C.byref(nb))
    if r < 0:
        return out[: nb.value].copy(), int(err.value)
    return out[:r].copy(), None


def translate_codon(codon):
    """encodings.rs:298-326; raises ValueError for lengths outside 1..3."""
    if isinstance(codon, str):
        codon = codon.encode()
    r = lib.orc_translate_codon(codon, len(codon))
    if r == 0:
        raise ValueError("Codon is invalid length: %d" % len(codon))
    return chr(r)


def seq_to_hashes_protein(seq, ksize_aa, moltype="protein", seed=42, keep_zeros=False):
    """SeqToHashes(is_protein=True): residues -> hashes in order (signature.rs:358-388).
    ksize_aa is in residues.  Raises ValueError for a DNA sketch (InvalidHashFunction)."""
    if isinstance(seq, str):
        seq = seq.encode()
    out = np.zeros(len(seq) + 1, dtype=np.uint64)
    n = lib.orc_seq_to_hashes_protein(seq, len(seq), 3 * ksize_aa, seed, HASH_FUNCTIONS[moltype],
                                      int(keep_zeros), out)
    if n < 0:
        raise ValueError("Invalid hash function")
    return out[:n].copy()


def seq_to_hashes_translate(seq, ksize_aa, moltype="protein", seed=42, keep_zeros=False):
    """SeqToHashes(is_protein=False) on a protein-family sketch: six-frame translation
    (signature.rs:307-357)."""
    if isinstance(seq, str):
        seq = seq.encode()
    out = np.zeros(2 * len(seq) + 4, dtype=np.uint64)
    n = lib.orc_seq_to_hashes_translate(seq, len(seq), 3 * ksize_aa, seed, HASH_FUNCTIONS[moltype],
                                        int(keep_zeros), out)
    return out[:n].copy()


class OracleMinHash:
    """Vec-backed KmerMinHash restatement (src/core/src/sketch/minhash.rs:41-702)."""

    def __init__(self, scaled=0, ksize=31, seed=42, track_abundance=False, num=0):
        self.ksize, self.seed, self.num = ksize, seed, num
        self.track = track_abundance
        self._p = lib.orc_mh_new(scaled, ksize, seed, int(track_abundance), num)

    def __del__(self):
        if getattr(self, "_p", None):
            lib.orc_mh_free(self._p)
            self._p = None

    def __len__(self):
        return int(lib.orc_mh_size(self._p))

    @property
    def max_hash(self):
        return int(lib.orc_mh_max_hash(self._p))

    def mins(self):
        n = len(self)
        if n == 0:
            return np.zeros(0, dtype=np.uint64)
        return np.ctypeslib.as_array(lib.orc_mh_mins(self._p), shape=(n,)).copy()

    def abunds(self):
        n = len(self)
        p = lib.orc_mh_abunds(self._p)
        if not p or n == 0:
            return np.zeros(0, dtype=np.uint64)
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def add_hash(self, h):
        lib.orc_mh_add_hash(self._p, h)

    def add_hash_with_abundance(self, h, a):
        lib.orc_mh_add_hash_with_abundance(self._p, h, a)

    def add_many(self, hs):
        hs = _u64(hs)
        lib.orc_mh_add_many(self._p, hs, len(hs))

    def remove_hash(self, h):
        lib.orc_mh_remove_hash(self._p, h)

    def add_sequence(self, seq, force=False):
        """Returns None, or the index of the first invalid window when force is False."""
        if isinstance(seq, str):
            seq = seq.encode()
        r = lib.orc_mh_add_sequence(self._p, seq, len(seq), int(force))
        return None if r == 0 else int(r - 1)

    def add_protein_family(self, seq, moltype, input_is_protein):
        """add_protein (input_is_protein) / add_sequence (translate) on a protein-family sketch;
        self.ksize must be the ABI value (3 x residues)."""
        if isinstance(seq, str):
            seq = seq.encode()
        r = lib.orc_mh_add_protein_family(self._p, seq, len(seq), HASH_FUNCTIONS[moltype],
                                          int(input_is_protein))
        if r < 0:
            raise ValueError("Invalid hash function")

    def merge(self, other):
        lib.orc_mh_merge(self._p, other._p)

    def md5sum(self):
        return md5sum(self.ksize, self.mins())


def md5sum(ksize, mins):
    mins = _u64(mins)
    buf = C.create_string_buffer(33)
    lib.orc_md5sum(ksize, mins, len(mins), buf)
    return buf.value.decode()


def count_common(a, b):
    a, b = _u64(a), _u64(b)
    return int(lib.orc_count_common(a, len(a), b, len(b)))


def intersection_size(a, b, num=0):
    a, b = _u64(a), _u64(b)
    c, u = C.c_uint64(0), C.c_uint64(0)
    if num:
        lib.orc_intersection_size_num(a, len(a), b, len(b), num, C.byref(c), C.byref(u))
    else:
        lib.orc_intersection_size(a, len(a), b, len(b), C.byref(c), C.byref(u))
    return int(c.value), int(u.value)


def jaccard(a, b, num=0):
    a, b = _u64(a), _u64(b)
    return float(lib.orc_jaccard(a, len(a), b, len(b), num))


def angular_similarity(a, aa, b, ba):
    a, aa, b, ba = _u64(a), _u64(aa), _u64(b), _u64(ba)
    return float(lib.orc_angular_similarity(a, aa, len(a), b, ba, len(b)))


def downsample(a, new_max_hash):
    a = _u64(a)
    return a[: lib.orc_downsample_count(a, len(a), new_max_hash)]


def to_csr(rows):
    offsets = np.zeros(len(rows) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(r) for r in rows])
    hashes = np.concatenate([_u64(r) for r in rows]) if len(rows) else np.zeros(0, np.uint64)
    return _u64(hashes), offsets


def compare_all_pairs(hashes, offsets, num=0, first_row=0, n_rows=None, nthreads=1):
    """compare_serial restated (src/sourmash/compare.py:14-64).  Rows outside
    [first_row, first_row+n_rows) are left as ones."""
    n = len(offsets) - 1
    if n_rows is None:
        n_rows = n - first_row
    out = np.ones((n, n), dtype=np.float64)
    lib.orc_compare_all_pairs(_u64(hashes), _u64(offsets), n, num, first_row, n_rows, out, nthreads)
    return out


def pairwise_common(hashes, offsets, first_row=0, n_rows=None, nthreads=1):
    n = len(offsets) - 1
    if n_rows is None:
        n_rows = n - first_row
    out = np.zeros((n, n), dtype=np.uint32)
    lib.orc_pairwise_common(_u64(hashes), _u64(offsets), n, first_row, n_rows, out, nthreads)
    return out


def one_vs_many(q, hashes, offsets, nthreads=1):
    q = _u64(q)
    n = len(offsets) - 1
    out = np.zeros(n, dtype=np.uint64)
    lib.orc_one_vs_many(q, len(q), _u64(hashes), _u64(offsets), n, out, nthreads)
    return out


def one_vs_many_bsearch(q, hashes, offsets, nthreads=1):
    "one_vs_many for a query much larger than the subjects: same counts, O(|S| log |Q|) per subject (oracle.c)"
    q = _u64(q)
    n = len(offsets) - 1
    out = np.zeros(n, dtype=np.uint64)
    lib.orc_one_vs_many_bsearch(q, len(q), _u64(hashes), _u64(offsets), n, out, nthreads)
    return out


def sketch_scaled(seq, k, max_hash, seed=42):
    if isinstance(seq, (bytes, bytearray, str)):
        seq = np.frombuffer(seq.encode() if isinstance(seq, str) else bytes(seq), dtype=np.uint8)
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    cap = max(len(seq), 1)
    out = np.zeros(cap, dtype=np.uint64)
    nk = C.c_uint64(0)
    n = lib.orc_sketch_scaled(seq, len(seq), k, seed, max_hash, out, cap, C.byref(nk))
    return out[:n].copy()


def sketch_batch(seqs, seq_off, k, max_hash, seed=42, nthreads=1, cap_per_seq=None):
    """OpenMP-over-genomes sketching; returns list of arrays."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    seq_off = _u64(seq_off)
    n = len(seq_off) - 1
    lens = np.diff(seq_off.astype(np.int64))
    if cap_per_seq is None:
        cap_per_seq = int(lens.max()) if n else 0
    out_off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(cap_per_seq)).astype(np.uint64)
    out = np.zeros(int(out_off[-1]), dtype=np.uint64)
    out_n = np.zeros(n, dtype=np.uint64)
    lib.orc_sketch_batch(seqs, seq_off, n, k, seed, max_hash, out, out_off, out_n, nthreads)
    return [out[int(out_off[i]): int(out_off[i]) + int(min(out_n[i], cap_per_seq))].copy()
            for i in range(n)]
