"""A .zip collection searched from HBM: ZipFileLinearIndex (parsed natively, one CSR upload per
scaled group) against LinearIndex over the same signatures loaded one object at a time, and against
the reference's known gather answer for the 12-genome fixture (tests/test_index_protocol.py:1057-1097)."""
import glob
import os

import pytest

from tests.conftest import GOLDEN
from tests.test_zip_collections import make_zip_deflated

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def smb():
    import sourmash_b200
    return sourmash_b200


def test_zip_index_search_prefetch_gather(smb, golden, tmp_path):
    from sourmash_b200.index import LinearIndex, ZipFileLinearIndex, gather
    paths = sorted(glob.glob(os.path.join(GOLDEN, "gather", "GCF*.sig")))
    members = [("signatures/%s" % os.path.basename(p), p) for p in paths]
    gz = {members[1][0], members[7][0]}
    query = smb.signature.load_one_signature_from_json(os.path.join(GOLDEN, "gather", "combined.sig"), ksize=21)
    for use_manifest in (True, False):
        zpath = make_zip_deflated(tmp_path / ("db%d.zip" % use_manifest), members, manifest=use_manifest, gz=gz)
        idx = ZipFileLinearIndex.load(zpath, use_manifest=use_manifest).select(ksize=21, moltype="DNA")
        lin = LinearIndex(idx.signatures())
        assert len(idx) == len(lin) == 12
        a = idx.search(query, threshold=0.0, do_containment=True)
        b = lin.search(query, threshold=0.0, do_containment=True)
        assert [(r.score, r.signature.md5sum()) for r in a] == [(r.score, r.signature.md5sum()) for r in b]
        assert len(a) == 12 and all(r.location == zpath for r in a)
        j = idx.search(query, threshold=0.0)
        assert [(r.score, r.signature.md5sum()) for r in j] == \
            [(r.score, r.signature.md5sum()) for r in lin.search(query, threshold=0.0)]
        pa = [(r.score, r.signature.md5sum()) for r in idx.prefetch(query, 50000)]
        assert pa and pa == [(r.score, r.signature.md5sum()) for r in lin.prefetch(query, 50000)]
        hits = list(gather(query, idx, threshold_bp=0))
        assert [[h.match.name.split()[0], h.intersect_size] for h in hits] == golden["meta"]["gather_k21_expected"]
        assert [(h.match.md5sum(), h.intersect_size) for h in hits] == \
            [(h.match.md5sum(), h.intersect_size) for h in gather(query, lin, threshold_bp=0)]


def test_compare_signature_files_from_zip(smb, tmp_path):
    "compare over an archive == compare over its members as plain files."
    import numpy as np
    from sourmash_b200.sigset import compare_signature_files
    paths = sorted(glob.glob(os.path.join(GOLDEN, "gather", "GCF*.sig")))
    members = [("signatures/%s" % os.path.basename(p), p) for p in paths]
    zpath = make_zip_deflated(tmp_path / "db.zip", members)
    m_zip, labels_zip = compare_signature_files([zpath], ksize=31)
    m_plain, labels_plain = compare_signature_files(paths, ksize=31)
    assert labels_zip == labels_plain and np.array_equal(m_zip, m_plain)
    assert m_zip.shape == (12, 12) and np.array_equal(np.diag(m_zip), np.ones(12))
