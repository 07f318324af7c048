"""Builds an EMULATED copy of the library for the CPU-only suite: the sources of sourmash_b200/csrc are copied
to a scratch directory, every kernel launch ``k<<<grid, block, smem, stream>>>(args)`` is rewritten into a call
of the SIMT emulator (simt.h: CTAs as cooperative fibers), ``__shared__`` declarations become static / emulator
storage, and the result is compiled with g++ against host stand-ins for the CUDA runtime (mock/mock_cudart.cpp)
and for the cub algorithms (mock/cub).  The C ABI, the host glue of capi.cu and the kernels are the product's
own code; only the "device" is the CPU.

TEST INFRASTRUCTURE ONLY.  The emulated library is written under the system temp directory, is never shipped or
loaded by the package (sourmash_b200/_lowlevel.py only knows libsourmash_b200.so and fails when it is missing);
tests/test_emulated_library.py loads it in a subprocess in place of the real library to run host glue that
otherwise needs a GPU -- in particular the paths that sit behind switches and have not run on a GPU yet.
"""
import os
import re
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "sourmash_b200", "csrc")
SOURCES = ["capi.cu", "sketch_kernels.cu", "compare_kernels.cu", "ingest.cu"]


def _match_paren(text, i):
    "index just past the parenthesis that closes the one opening at text[i]"
    depth = 0
    for j in range(i, len(text)):
        if text[j] == "(":
            depth += 1
        elif text[j] == ")":
            depth -= 1
            if depth == 0:
                return j + 1
    raise ValueError("unbalanced parentheses")


def _split_top(text):
    "split on commas that are not nested in (), <> or []"
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


_LAUNCH = re.compile(r"([A-Za-z_][A-Za-z0-9_]*(?:<[^<>;()]*>)?)\s*<<<")


def rewrite_launches(text):
    out, at = "", 0
    while True:
        m = _LAUNCH.search(text, at)
        if not m:
            return out + text[at:]
        end_cfg = text.index(">>>", m.end())
        cfg = _split_top(text[m.end():end_cfg])
        while len(cfg) < 4:
            cfg.append("0")
        grid, block, smem, _stream = cfg[:4]
        i = end_cfg + 3
        while text[i].isspace():
            i += 1
        assert text[i] == "(", text[m.start():i + 20]
        j = _match_paren(text, i)
        args = text[i + 1:j - 1]
        out += text[at:m.start()]
        out += "smb_emu::launch(%s, %s, (size_t)(%s), [&] { %s(%s); })" % (grid, block, smem, m.group(1), args)
        at = j


def rewrite_shared(text):
    text = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([A-Za-z_][A-Za-z0-9_ ]*?)\s+([A-Za-z_][A-Za-z0-9_]*)\[\];",
                  r"SMB_DYN_SHARED(\1, \2);", text)
    return re.sub(r"\b__shared__\b", "static", text)


def emulated_sources(dst):
    os.makedirs(dst, exist_ok=True)
    for name in os.listdir(CSRC):
        src = os.path.join(CSRC, name)
        if not os.path.isfile(src):
            continue
        with open(src) as fh:
            text = fh.read()
        if name.endswith(".cu"):
            text = rewrite_shared(rewrite_launches(text))
            name = name[:-3] + ".cpp"
        elif name.endswith(".cuh"):
            text = rewrite_shared(text) if "SMB_SHARED" not in text else text
        with open(os.path.join(dst, name), "w") as fh:
            fh.write(text)


def _newest_input():
    t = 0.0
    for d in (CSRC, HERE, os.path.join(HERE, "mock"), os.path.join(ROOT, "include")):
        for base, _dirs, files in os.walk(d):
            for f in files:
                t = max(t, os.path.getmtime(os.path.join(base, f)))
    return t


def build(verbose=False):
    "Returns the path of the emulated shared library, (re)building it when an input is newer."
    import fcntl
    with open(os.path.join(tempfile.gettempdir(), "smb_emul_lib.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)                  # concurrent test processes build it once
        return _build_locked(verbose)


def _build_locked(verbose):
    # SMB_EMUL_ASAN=1: a second build with AddressSanitizer + UBSan for the host code (parsers, glue); load it with
    # LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 (tests/tools/fuzz_ingest.py does)
    asan = os.environ.get("SMB_EMUL_ASAN") == "1"
    tsan = not asan and os.environ.get("SMB_EMUL_TSAN") == "1"       # ThreadSanitizer: the multi-threaded file readers
    top = os.path.join(tempfile.gettempdir(), "smb_emul_lib_asan" if asan else "smb_emul_lib_tsan" if tsan else "smb_emul_lib")
    lib = os.path.join(top, "libsourmash_b200_emul.so")
    if os.path.exists(lib) and os.path.getmtime(lib) >= _newest_input():
        return lib
    work = os.path.join(top, "sourmash_b200", "csrc")          # keeps the ../../include/sourmash_b200.h relation
    shutil.rmtree(top, ignore_errors=True)
    emulated_sources(work)
    os.makedirs(os.path.join(top, "include"), exist_ok=True)            # csrc includes ../../include/sourmash_b200.h
    shutil.copy(os.path.join(ROOT, "include", "sourmash_b200.h"), os.path.join(top, "include", "sourmash_b200.h"))
    flags = ["-O1", "-g", "-std=c++17", "-fPIC", "-w", "-DSMB_SIMT_EMUL=1", "-include", os.path.join(HERE, "simt.h"),
             "-I", os.path.join(HERE, "mock"), "-I", "/usr/local/cuda/include", "-I", work]
    san = ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fno-sanitize-recover=undefined"] if asan else []
    if tsan:
        san = ["-fsanitize=thread", "-fno-omit-frame-pointer"]
    flags += san
    objs, procs = [], []
    for src in [s[:-3] + ".cpp" for s in SOURCES]:
        obj = os.path.join(work, src[:-4] + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen(["g++"] + flags + ["-c", os.path.join(work, src), "-o", obj],
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    mock_obj = os.path.join(work, "mock_cudart.o")
    procs.append(subprocess.Popen(["g++", "-O1", "-std=c++17", "-fPIC", "-w", "-I", "/usr/local/cuda/include", "-c",
                                   os.path.join(HERE, "mock", "mock_cudart.cpp"), "-o", mock_obj],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    failed = ""
    for p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed += out
    if failed:
        raise RuntimeError("emulated build failed:\n" + failed[-6000:])
    subprocess.check_call(["g++", "-shared", "-o", lib] + san + objs + [mock_obj, "-lz", "-ldl", "-lpthread"])
    if verbose:
        print("built", lib)
    return lib


if __name__ == "__main__":
    print(build(verbose=True))
