"""Build recipe for libsourmash_b200.so (nvcc, sm_100a only).  Used by __graft_entry__.build()."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsourmash_b200.so")
SOURCES = ["capi.cu", "sketch_kernels.cu", "compare_kernels.cu", "ingest.cu"]
HEADERS = ["common.cuh", "kernels.h", "md5.h", "kmer_roll.cuh", "split_table.cuh", "aa_kmers.cuh", "ingest.h",
           "join_walk.cuh", "join_stripe.cuh", "range_search.cuh", "db_index.cuh", "experimental_kernels.cuh", "sketch_device.cuh", "search_kernels.cuh", "tile_kernels.cuh", "pair_kernels.cuh", "zipread.h", os.path.join("..", "..", "include", "sourmash_b200.h")]
OBJ_DIR = os.path.join(CSRC, "build")                 # git-ignored; only the linked .so ships
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newest_header():
    return max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)


def _stale_objects():
    hdr = _newest_header()
    out = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr, os.path.getmtime(os.path.join(CSRC, src))):
            out.append((src, obj))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """One nvcc -c per source, side by side, then one link: a change to one file recompiles that file."""
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    todo = [(s, os.path.join(OBJ_DIR, s[:-3] + ".o")) for s in SOURCES] if force else _stale_objects()
    procs = []
    for src, obj in todo:
        cmd = [nvcc_path()] + NVCC_FLAGS + ["-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, cwd=CSRC)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-shared", "-o", LIB] + \
        [os.path.join(OBJ_DIR, s[:-3] + ".o") for s in SOURCES] + ["-lz", "-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
