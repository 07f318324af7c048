"""The reference's OWN Python -- object model, index classes, search / gather helpers, command line -- over this
library's C ABI.

/root/reference/src/sourmash is loaded IN PLACE (symlinks, nothing is copied) as a package `sourmash` whose
`_lowlevel` is a cffi ABI-mode binding built from the REFERENCE header /root/reference/include/sourmash.h --
exactly what maturin generates for the Rust cdylib (pyproject.toml:138-155) -- but dlopen()s libsourmash_b200
instead.  The reference's own test modules (unmodified, read in place with their conftest.py and test-data) are
then run by pytest in a subprocess: the object model (test_minhash.py, test_jaccard.py, test__minhash_hypothesis.py,
test_signature.py, test_sketchcomparison.py), the reference's compare.py loops (test_compare.py), its search / index
classes (test_search.py, test_index.py, test_index_protocol.py, test_api.py, test_manifest.py, test_picklist.py) and
its command line (`sourmash sketch ...`: test_sourmash_sketch.py; the CLI is run in-process through the entry point
of a dist-info written next to the package, as sourmash_tst_utils.runscript expects).

No GPU here and no /root/reference on the GPU box, so the "device" is the emulated build of the library
(tests/host_emul/emul_lib.py: the product's capi.cu host glue and kernels compiled for the CPU, kernel
launches executed by the SIMT emulator).  What this proves is the BOUNDARY: every symbol, struct layout,
ownership rule, error code and message the reference's Python touches on these paths.  The kernels' results on
real hardware are pinned separately by the -m gpu tests against the same oracle / golden vectors.

Ours, and tiny: stand-ins for three third-party packages that are not installed (`deprecation`: a decorator that warns;
`screed`: rc + a FASTA / FASTQ(.gz/.bz2) record iterator; `matplotlib`: conftest.py touches pyplot.rcParams), and
placeholders for the reference modules that bind OUT-OF-SCOPE engines: sbt.py / sbtmh.py / nodegraph.py
(nodegraph_* FFI), index/revindex.py (revindex_* FFI), index/sqlite_index.py (needs `bitstring`).  Their loaders answer
"not this format", so the reference's loader chain (save_load.py) walks past them.  REFERENCE_DESELECT lists the
reference tests that are not run, each with its reason; everything else in the modules must pass."""
import os
import subprocess
import sys
import textwrap

import pytest

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "sourmash")),
                                reason="needs the reference checkout (this container only)")

LOWLEVEL = '''
"""cffi ABI-mode binding generated from the REFERENCE header, loading libsourmash_b200 (emulated build)."""
import os, re
import cffi
with open({header!r}) as fh:
    text = fh.read()
text = re.sub(r"/\\*.*?\\*/", "", text, flags=re.S)
lines = [l for l in text.splitlines() if not l.strip().startswith("#")]       # build.rs:50-74 strips them for cffi too
ffi = cffi.FFI()
ffi.cdef("\\n".join(lines))
lib = ffi.dlopen({lib!r})
'''

INIT = '''
"""Package __init__: the imports of /root/reference/src/sourmash/__init__.py in its order (the modules import each other
in circles that only resolve in this order), without its deprecated wrapper functions."""
from ._lowlevel import ffi, lib
ffi.init_once(lib.sourmash_init, "init")
VERSION = "4.9.9+libsourmash_b200"
from .minhash import MinHash, get_minhash_default_seed, get_minhash_max_hash
DEFAULT_SEED = get_minhash_default_seed()
MAX_HASH = get_minhash_max_hash()
from .signature import load_signatures_from_json, load_one_signature_from_json, SourmashSignature, save_signatures_to_json
load_signatures, load_one_signature, save_signatures = load_signatures_from_json, load_one_signature_from_json, save_signatures_to_json
from .sbtmh import load_sbt_index, search_sbt_index, create_sbt_index
from . import lca
from . import tax
from . import sbt
from . import sbtmh
from . import sbt_storage
from . import signature
from . import sig
from . import cli
from . import commands
from .sourmash_args import load_file_as_index
from .sourmash_args import load_file_as_signatures
'''

# third-party packages that are not installed here
STUBS = {
    "deprecation.py": textwrap.dedent("""
        import functools, warnings
        def deprecated(*a, **k):                                   # the PyPI package warns once the version is reached
            def deco(f):
                @functools.wraps(f)
                def wrapper(*args, **kwargs):
                    warnings.warn(f.__name__ + " is deprecated", DeprecationWarning, stacklevel=2)
                    return f(*args, **kwargs)
                return wrapper
            return deco
        """),
    "matplotlib/__init__.py": "def use(*a, **k): pass\n",      # tests/test_sourmash.py:31-33; `sourmash plot` itself is deselected
    "matplotlib/pyplot.py": "rcParams = {}\n",                  # tests/conftest.py:7-9
    "screed/__init__.py": textwrap.dedent('''
        """rc() and a FASTA / FASTQ record iterator (plain, .gz, .bz2, '-' = stdin): what the reference uses of screed"""
        import bz2, gzip, io, sys
        __version__ = "0+stand-in"                                  # `sourmash info -v` prints it
        _C = str.maketrans("ACGTNacgtn", "TGCANtgcan")
        def rc(s):
            return s.translate(_C)[::-1]
        class Record(dict):
            def __init__(self, name=None, sequence=None, **kw):
                super().__init__(name=name, sequence=sequence, **kw)
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)
            def __len__(self):
                return len(self["sequence"])
        def _text(path):
            if path in ("-", "/dev/stdin"):
                data = sys.stdin.buffer.read() if hasattr(sys.stdin, "buffer") else sys.stdin.read().encode()
            else:
                with io.open(path, "rb") as fh:
                    data = fh.read()
            if data[:2] == b"\\x1f\\x8b":
                data = gzip.decompress(data)
            elif data[:3] == b"BZh":
                data = bz2.decompress(data)
            return data.decode("utf-8", "replace")
        def _parse(text):
            lines = [l.rstrip() for l in text.splitlines()]
            i = 0
            while i < len(lines) and not lines[i]:
                i += 1
            if i >= len(lines):
                return                                              # no records
            if lines[i][0] == ">":
                name, chunks = None, []
                for line in lines[i:]:
                    if line.startswith(">"):
                        if name is not None:
                            yield Record(name, "".join(chunks))
                        name, chunks = line[1:], []
                    elif line:
                        chunks.append(line)
                if name is not None:
                    yield Record(name, "".join(chunks))
            elif lines[i][0] == "@":
                while i < len(lines):
                    if not lines[i]:
                        i += 1
                        continue
                    if lines[i][0] != "@" or i + 1 >= len(lines):
                        raise ValueError("not a FASTQ record")
                    yield Record(lines[i][1:], lines[i + 1], quality=lines[i + 3] if i + 3 < len(lines) else "")
                    i += 4
            else:
                raise ValueError("unknown sequence file format")
        class _Records:                                             # an iterator that is also a context manager, like screed's
            def __init__(self, path):
                self._it = _parse(_text(str(path)))
                try:
                    self._first = [next(self._it)]                  # format errors surface at open()
                except StopIteration:
                    self._first = []
            def __bool__(self):                                     # `if not screed_iter:` -- no records (command_sketch.py:698)
                return bool(self._first)
            def __iter__(self):
                return self
            def __next__(self):
                if self._first:
                    return self._first.pop()
                return next(self._it)
            def __enter__(self):
                return self
            def __exit__(self, *a):
                return False
            def close(self):
                pass
        def open(path, *a, **k):
            return _Records(path)
        '''),
}

# Reference modules that bind OUT-OF-SCOPE engines: importable names whose loaders say "not this format"
OUT_OF_SCOPE_STUBS = {
    "sbt.py": "class SBT: pass\nclass Leaf: pass\nclass Node: pass\nclass GraphFactory:\n    def __init__(self, *a, **k): pass\n",
    "sbtmh.py": textwrap.dedent("""
        def load_sbt_index(*a, **k): raise ValueError('not an SBT (SBT needs the nodegraph FFI: out of scope here)')
        def create_sbt_index(*a, **k): raise NotImplementedError('SBT: out of scope here')
        def search_sbt_index(*a, **k): raise NotImplementedError('SBT: out of scope here')
        class SigLeaf: pass
        class LocalizedSBT: pass
        """),
    "nodegraph.py": textwrap.dedent("""
        class Nodegraph: pass
        def extract_nodegraph_info(*a): raise NotImplementedError
        def calc_expected_collisions(*a, **k): raise NotImplementedError
        """),
    "index/revindex.py": "class RevIndex: pass\n",               # the Rust RevIndex (revindex_* symbols)
    "index/sqlite_index.py": textwrap.dedent('''
        """SqliteIndex and friends need `bitstring` (not installed): every loader says 'not a sqlite database'"""
        class _NotSqlite:
            @classmethod
            def load(cls, *a, **k): raise ValueError("not a sqlite database (SqliteIndex is out of scope here)")
            load_from_filename = load
            create = load
            def __init__(self, *a, **k): raise ValueError("SqliteIndex is out of scope here")
        class SqliteIndex(_NotSqlite): pass
        class SqliteCollectionManifest(_NotSqlite): pass
        class LCA_SqliteDatabase(_NotSqlite): pass
        def load_sqlite_index(*a, **k): return None
        '''),
}

# the reference package, loaded in place
IN_PLACE = ("minhash.py", "signature.py", "utils.py", "exceptions.py", "distance_utils.py", "logging.py", "compare.py",
            "sketchcomparison.py", "np_utils.py", "search.py", "manifest.py", "picklist.py", "sbt_storage.py", "sourmash_args.py",
            "save_load.py", "plugins.py", "sqlite_utils.py", "commands.py", "command_sketch.py", "command_compute.py", "fig.py",
            "__main__.py", "cli", "sig", "lca", "tax")
TEST_FILES = ("conftest.py", "sourmash_tst_utils.py", "test-data", "test_minhash.py", "test_jaccard.py",
              "test__minhash_hypothesis.py", "test_signature.py", "test_compare.py", "test_sketchcomparison.py", "test_search.py",
              "test_index_protocol.py", "test_index.py", "test_api.py", "test_manifest.py", "test_picklist.py",
              "test_sourmash_sketch.py", "test_sourmash.py", "test_prefetch.py", "test_sourmash_compute.py", "test_cmd_signature.py")


def _stub_package(tmp, lib_path):
    site = os.path.join(tmp, "site")                               # not next to the tests: scriptpath() looks for ../sourmash
    pkg = os.path.join(site, "sourmash")
    os.makedirs(os.path.join(pkg, "index"))
    for name in IN_PLACE:
        os.symlink(os.path.join(REF, "src", "sourmash", name), os.path.join(pkg, name))
    os.symlink(os.path.join(REF, "src", "sourmash", "index", "__init__.py"), os.path.join(pkg, "index", "__init__.py"))
    for rel, text in OUT_OF_SCOPE_STUBS.items():
        with open(os.path.join(pkg, rel), "w") as fh:
            fh.write(text)
    with open(os.path.join(pkg, "_lowlevel.py"), "w") as fh:
        fh.write(LOWLEVEL.format(header=os.path.join(REF, "include", "sourmash.h"), lib=lib_path))
    with open(os.path.join(pkg, "__init__.py"), "w") as fh:
        fh.write(INIT)
    for rel, text in STUBS.items():
        path = os.path.join(site, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as fh:
            fh.write(text)
    # what `pip install` would leave behind: the version and the console script that sourmash_tst_utils.runscript looks up
    info = os.path.join(site, "sourmash-4.9.9.dist-info")
    os.makedirs(info)
    with open(os.path.join(info, "METADATA"), "w") as fh:
        fh.write("Metadata-Version: 2.1\nName: sourmash\nVersion: 4.9.9\n")
    with open(os.path.join(info, "entry_points.txt"), "w") as fh:
        fh.write("[console_scripts]\nsourmash = sourmash.__main__:main\n")
    os.makedirs(os.path.join(tmp, "bin"))                          # ... and the script itself: runscript falls back to exec()ing it
    with open(os.path.join(tmp, "bin", "sourmash"), "w") as fh:   # when the entry point raised ValueError (sourmash_tst_utils.py:49-71)
        fh.write("import sys\nfrom sourmash.__main__ import main\nif __name__ == '__main__':\n    sys.exit(main())\n")
    # the reference's test modules, conftest and helpers, read in place
    tests = os.path.join(tmp, "reftests")
    os.makedirs(tests)
    for name in TEST_FILES:
        os.symlink(os.path.join(REF, "tests", name), os.path.join(tests, name))
    return tests


def _run_reference_tests(tmp_path, modules, extra=()):
    sys.path.insert(0, os.path.join(HERE, "host_emul"))
    try:
        import emul_lib
    finally:
        sys.path.pop(0)
    lib_path = emul_lib.build()
    tmp = str(tmp_path)
    tests = _stub_package(tmp, lib_path)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(tmp, "site"), tests]), PYTHONDONTWRITEBYTECODE="1",
               PATH=os.path.join(tmp, "bin") + os.pathsep + os.environ.get("PATH", ""))
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", "--rootdir", tests, "-o", "python_files=test_*.py",
           "--no-header", "-rN"] + _workers() + list(extra) + [os.path.join(tests, m) for m in modules]
    return subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=tmp, timeout=3000)


def _workers():
    "pytest-xdist workers for the subprocess, when the plugin is there (the command-line suites take minutes on one core)"
    try:
        import xdist  # noqa: F401
    except ImportError:
        return []
    return ["-n", str(min(8, os.cpu_count() or 1))]


def _counts(stdout):
    import re
    tail = stdout.strip().splitlines()[-1] if stdout.strip() else ""
    out = {k: 0 for k in ("passed", "failed", "error", "errors", "skipped", "deselected", "xfailed")}
    for num, word in re.findall(r"(\d+) (\w+)", tail):
        if word in out:
            out[word] = int(num)
    return out, tail


def _deselect(modules):
    out = []
    for mod in modules:
        for d in REFERENCE_DESELECT.get(mod, ()):
            out += ["--deselect", "%s::%s" % (mod, d)]                 # ids are relative to --rootdir = the directory of the test modules
    return out


def _check(r, at_least):
    counts, tail = _counts(r.stdout)
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-3000:]
    assert counts["failed"] == 0 and counts["error"] + counts["errors"] == 0 and counts["passed"] >= at_least, tail


@pytest.mark.timeout(3600)
def test_reference_minhash_tests_pass_over_this_abi(tmp_path):
    """/root/reference/tests/test_minhash.py (199 test functions, 311 cases with its fixtures), unmodified."""
    mods = ["test_minhash.py"]
    _check(_run_reference_tests(tmp_path, mods, _deselect(mods)), 300)


@pytest.mark.timeout(1800)
def test_reference_jaccard_and_hypothesis_tests_pass_over_this_abi(tmp_path):
    "tests/test_jaccard.py (real-data Jaccard KATs, downsampling) and tests/test__minhash_hypothesis.py, unmodified"
    mods = ["test_jaccard.py", "test__minhash_hypothesis.py"]
    _check(_run_reference_tests(tmp_path, mods, _deselect(mods)), 15)


@pytest.mark.timeout(1800)
def test_reference_signature_compare_and_sketchcomparison_tests_pass_over_this_abi(tmp_path):
    """tests/test_signature.py (48 functions: SourmashSignature over signature_* / signatures_load_* / save), test_compare.py
    (the reference's compare_serial*, compare_parallel and compare_all_pairs loops calling this ABI pair by pair) and
    test_sketchcomparison.py (FracMinHashComparison / NumMinHashComparison: containment, ANI, downsampling), unmodified."""
    mods = ["test_signature.py", "test_compare.py", "test_sketchcomparison.py"]
    _check(_run_reference_tests(tmp_path, mods, _deselect(mods)), 140)


# test_index_protocol.py parametrises every test over 11 Index builders and 3 CounterGather flavours, test_index.py has tests
# per class; the ones built on SURVEY section 2's OUT-OF-SCOPE engines are not run: SBT (needs the nodegraph FFI), SqliteIndex
# (needs `bitstring`), RevIndex (revindex_* FFI).  What runs: LinearIndex, LazyLinearIndex, ZipFileLinearIndex (over this
# library's zipstorage_*), MultiIndex, StandaloneManifestIndex, LCA_Database (the reference's pure-Python class, JSON form),
# CounterGather, CounterGather_LinearIndex, CounterGather_LCA.
NOT_SBT_SQLITE_REVINDEX = "not sbt and not SBT and not sql and not Sql and not revindex and not RevIndex"


@pytest.mark.timeout(1800)
def test_reference_search_and_index_protocol_tests_pass_over_this_abi(tmp_path):
    """tests/test_search.py (39 functions: the reference's JaccardSearch / search / gather helpers and LinearIndex.find loops)
    and tests/test_index_protocol.py -- the reference's conformance suite for Index and CounterGather classes -- for every
    class listed above, unmodified; the reference's own index/__init__.py, search.py, manifest.py, picklist.py, lca/,
    sbt_storage.py, save_load.py, sourmash_args.py loaded in place.  Every count_common / intersection / downsample / zip
    read they make goes through this library."""
    mods = ["test_search.py", "test_index_protocol.py"]
    _check(_run_reference_tests(tmp_path, mods, ["-k", NOT_SBT_SQLITE_REVINDEX] + _deselect(mods)), 225)


@pytest.mark.timeout(1800)
def test_reference_index_tests_pass_over_this_abi(tmp_path):
    """tests/test_index.py: LinearIndex, LazyLinearIndex, ZipFileLinearIndex (manifests, select, traverse), MultiIndex
    (directories, pathlists), StandaloneManifestIndex, CounterGather -- the reference's classes over this library, zip reads
    through zipstorage_* -- plus the public API (test_api.py), manifests and picklists."""
    mods = ["test_index.py", "test_api.py", "test_manifest.py", "test_picklist.py"]
    deselect = ["-k", NOT_SBT_SQLITE_REVINDEX + " and not simple_index"] + _deselect(mods)
    _check(_run_reference_tests(tmp_path, mods, deselect), 110)


@pytest.mark.timeout(1800)
def test_reference_sketch_command_line_tests_pass_over_this_abi(tmp_path):
    """tests/test_sourmash_sketch.py: `sourmash sketch dna | protein | translate | fromfile` -- the reference's CLI, its
    command_sketch.py (_compute_individual / _compute_merged), ComputeParameters and signature_add_sequence /
    signature_add_protein over this library, in-process, outputs to .sig / .sig.gz / .zip / directories -- including the
    known-good check against a signature made "another way" with mmh3 (test_sourmash_sketch.py:1303-1321)."""
    mods = ["test_sourmash_sketch.py"]
    _check(_run_reference_tests(tmp_path, mods, ["-k", "not sqldb"] + _deselect(mods)), 115)


# Why a test of the big command-line suites may fail here: it needs an engine SURVEY section 2 puts out of scope.  Each marker is
# text that only the placeholder modules above (or a missing third-party package) put into a failure report.
OUT_OF_SCOPE_MARKERS = {
    "SBT": ("SBT: out of scope here", "not an SBT", "'SBT'", ".sbt.zip", ".sbt.json", "SigLeaf", "Nodegraph", "sbt_combine"),
    "sqlite": ("SqliteIndex is out of scope", "not a sqlite database", ".sqldb", "sqlite_index", "SqliteCollectionManifest"),
    "plot": ("pylab", "scipy.cluster"),
    "repository file": ("CITATION.cff",),
}


def _failed_sections(stdout):
    "name -> report text of every failed / errored test in a `--tb=short` pytest report"
    import re
    parts = re.split(r"\n_{3,} (?:ERROR at \w+ of )?(test_[^\s]+) _{3,}\n", stdout)
    return {parts[i]: parts[i + 1] for i in range(1, len(parts) - 1, 2)}


@pytest.mark.timeout(3000)
def test_reference_command_line_suites_pass_over_this_abi(tmp_path):
    """The reference's command line end to end over this library: tests/test_sourmash.py (compare, search, gather, prefetch,
    multigather, sig loading ...: SURVEY 3.2-3.4's call stacks from the top), test_prefetch.py, test_sourmash_compute.py and
    test_cmd_signature.py (`sourmash sig merge / intersect / subtract / downsample / flatten / inflate ...`: out of scope as a
    feature, but every one of them is MinHash arithmetic through this ABI).  These suites also index SBTs, plot matrices and
    write sqlite databases; a test may fail only if its report carries one of OUT_OF_SCOPE_MARKERS, and the numbers of tests
    that pass are pinned from below."""
    mods = ["test_sourmash.py", "test_prefetch.py", "test_sourmash_compute.py", "test_cmd_signature.py"]
    r = _run_reference_tests(tmp_path, mods, ["--maxfail=100000", "--tb=short", "-rN"] + _deselect(mods))
    counts, tail = _counts(r.stdout)
    unexplained = {name: text[-1500:] for name, text in _failed_sections(r.stdout).items()
                   if not any(m in text or m in name for ms in OUT_OF_SCOPE_MARKERS.values() for m in ms)}
    assert not unexplained, "\n\n".join("%s\n%s" % kv for kv in list(unexplained.items())[:5])
    assert counts["passed"] >= 640, tail                           # 645 when written: 275 + 100 (prefetch, compute) + 270 (sig)
    assert counts["failed"] + counts["error"] + counts["errors"] <= 230, tail


# Reference tests that are NOT run, each with the reason; everything else in the modules must pass.
REFERENCE_DESELECT = {
    "test_minhash.py": [],
    "test_jaccard.py": [],
    "test__minhash_hypothesis.py": [],
    "test_signature.py": [],
    "test_compare.py": [],
    "test_sketchcomparison.py": [],
    "test_search.py": [],
    "test_index_protocol.py": [],
    "test_index.py": ["test_index_same_md5sum_fsstorage",            # `sourmash index`: builds an SBT
                      "test_standalone_manifest_load_from_dir"],     # a third of the manifest's rows live in all.sbt.zip
    "test_api.py": ["test_load_index_1",                 # loads an SBT
                    "test_load_and_search_sbt_api"],     # "
    "test_manifest.py": [],
    "test_picklist.py": [],
    "test_sourmash_sketch.py": [],
    "test_cmd_signature.py": ["test_sig_cat_4_filelist_with_dbs", "test_sig_cat_5_from_file",   # the file list names v6.sbt.zip (an SBT)
                              "test_sig_describe_3_manifest_works"],                           # scaled/mf.csv: a third of its rows are in all.sbt.zip
}
