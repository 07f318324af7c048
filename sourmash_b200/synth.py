"""Deterministic synthetic workloads, shared by tests and bench.py."""
import numpy as np

MAX_HASH_1000 = 18446744073709552


def synth_genome(length, seed, n_every=0, lower=False):
    """Uniform ACGT bytes from PCG64(seed) (config 2); optional N every n_every bases."""
    rng = np.random.Generator(np.random.PCG64(seed))
    codes = rng.integers(0, 4, size=length, dtype=np.uint8)
    lut = np.frombuffer(b"acgt" if lower else b"ACGT", dtype=np.uint8)
    seq = lut[codes]
    if n_every:
        seq = seq.copy()
        seq[n_every - 1::n_every] = ord("N")
    return seq


def synth_sketches(n, mean=5000, sd=500, lo=3000, hi=7000, n_families=100, pool=6000,
                   max_hash=MAX_HASH_1000, seed=0):
    """Config 3: n sketches in families sharing a hash pool.  Returns (hashes, offsets) CSR."""
    rng0 = np.random.Generator(np.random.PCG64(12345 + seed))
    sizes = np.clip(rng0.normal(mean, sd, size=n).round().astype(np.int64), lo, hi)
    pools = []
    for f in range(n_families):
        r = np.random.Generator(np.random.PCG64(2000 + f + 1000003 * seed))
        pools.append(np.unique(r.integers(1, max_hash, size=pool, dtype=np.uint64, endpoint=True)))
    rows = []
    for i in range(n):
        r = np.random.Generator(np.random.PCG64(3000 + i + 1000003 * seed))
        p = pools[i % n_families]
        frac = r.uniform(0.5, 0.95)
        take = min(int(frac * len(p)), int(sizes[i]))
        shared = r.choice(p, size=take, replace=False)
        extra = r.integers(1, max_hash, size=max(int(sizes[i]) - take, 0), dtype=np.uint64, endpoint=True)
        rows.append(np.unique(np.concatenate([shared, extra])))
    offsets = np.zeros(n + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(r) for r in rows])
    return np.concatenate(rows), offsets


def rows_of(hashes, offsets):
    return [hashes[int(offsets[i]):int(offsets[i + 1])] for i in range(len(offsets) - 1)]
