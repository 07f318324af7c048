"""Collection manifests: the per-sketch metadata table of a .zip collection.

Same rows, CSV format and selection rules as the reference's ``CollectionManifest``
(src/sourmash/manifest.py:15-390); what is added is ``from_signature_set``: the table of a
natively parsed collection comes out of the parser's metadata arrays instead of one Python
``make_manifest_row`` per loaded object.
"""
import ast
import csv
import itertools

_VERSION_LINE = "# SOURMASH-MANIFEST-VERSION: "


def _check_select_parameters(**kw):
    "Type checks of Index.select arguments (src/sourmash/index/__init__.py:1229-1272)."
    unknown = set(kw) - {"ksize", "num", "moltype", "scaled", "abund", "picklist", "containment"}
    if unknown:
        raise ValueError(f"unknown 'select' parameters: {unknown}")
    for key in ("ksize", "scaled", "num"):
        v = kw.get(key)
        if v is not None and not isinstance(v, int):
            raise ValueError(f"{key} value '{v}' must be an integer, is: {type(v)}")
    moltype = kw.get("moltype")
    if moltype is not None and moltype not in ["DNA", "protein", "dayhoff", "hp"]:
        raise ValueError(f"unknown moltype: {moltype}")
    for key in ("containment", "abund"):
        v = kw.get(key)
        if v is not None and not isinstance(v, bool):
            raise ValueError(f"{key} value '{v}' must be a bool, is: {type(v)}")


class CollectionManifest:
    "In-memory manifest: a list of row dictionaries (manifest.py:245-390)."

    required_keys = ("internal_location", "md5", "md5short", "ksize", "moltype", "num", "scaled",
                     "n_hashes", "with_abundance", "name", "filename")

    def __init__(self, rows=()):
        self.rows = []
        self._md5_set = set()
        self._add_rows(rows)

    def _add_rows(self, rows):
        for row in rows:
            self.rows.append(row)
            self._md5_set.add(row["md5"])

    def add_row(self, row):
        self._add_rows([row])

    # -- construction ----------------------------------------------------------------------
    @classmethod
    def load_from_csv(cls, fp):
        first = fp.readline().rstrip()
        if not first.startswith(_VERSION_LINE):
            raise ValueError("manifest is missing version header")
        version = first[len(_VERSION_LINE):]
        if float(version) != 1.0:
            raise ValueError(f"unknown manifest version number {version}")
        reader = csv.DictReader(fp)
        if not reader.fieldnames:
            raise ValueError("missing column headers in manifest")
        for key in cls.required_keys:
            if key not in reader.fieldnames:
                raise ValueError(f"missing column '{key}' in manifest.")
        rows = []
        for row in reader:
            for key in ("num", "scaled", "ksize", "n_hashes"):
                row[key] = int(row[key])
            row["with_abundance"] = bool(ast.literal_eval(str(row["with_abundance"])))
            row["signature"] = None
            rows.append(row)
        return cls(rows)

    @classmethod
    def make_manifest_row(cls, ss, location, *, include_signature=True):
        mh = ss.minhash
        md5 = ss.md5sum()
        row = {"internal_location": location, "md5": md5, "md5short": md5[:8], "ksize": int(mh.ksize),
               "moltype": mh.moltype, "num": int(mh.num), "scaled": int(mh.scaled), "n_hashes": len(mh),
               "with_abundance": mh.track_abundance, "name": ss.name, "filename": ss.filename}
        if include_signature:
            row["signature"] = ss
        return row

    @classmethod
    def create_manifest(cls, locations_iter, *, include_signature=True):
        return cls(cls.make_manifest_row(ss, loc, include_signature=include_signature)
                   for ss, loc in locations_iter)

    @classmethod
    def from_signature_set(cls, sigset, rows=None, md5s=None):
        """Rows for the sketches of a ``sigset.SignatureSet`` (all, or the listed ones), from its
        metadata arrays.  ``md5s``: identities already known (parallel to ``rows``)."""
        idx = range(len(sigset)) if rows is None else [int(r) for r in rows]
        scaled = sigset.python_scaled()
        out = []
        for pos, i in enumerate(idx):
            moltype = sigset.moltype(i)
            md5 = md5s[pos] if md5s is not None else sigset.md5sum(i)
            ksize = int(sigset.ksize[i])
            out.append({"internal_location": sigset.location(i), "md5": md5, "md5short": md5[:8],
                        "ksize": ksize if moltype == "DNA" else ksize // 3, "moltype": moltype,
                        "num": int(sigset.num[i]), "scaled": int(scaled[i]), "n_hashes": int(sigset.n_mins[i]),
                        "with_abundance": bool(sigset.has_abund[i]), "name": sigset.name(i),
                        "filename": sigset.filename(i), "signature": None})
        return cls(out)

    # -- output ----------------------------------------------------------------------------
    @classmethod
    def write_csv_header(cls, fp):
        fp.write(_VERSION_LINE + "1.0\n")
        csv.DictWriter(fp, fieldnames=cls.required_keys).writeheader()

    def write_to_csv(self, fp, write_header=False):
        w = csv.DictWriter(fp, fieldnames=self.required_keys, extrasaction="ignore")
        if write_header:
            self.write_csv_header(fp)
        for row in self.rows:
            w.writerow({k: v for k, v in row.items() if k != "signature"})

    # -- queries ---------------------------------------------------------------------------
    def _select(self, *, ksize=None, moltype=None, scaled=0, num=0, containment=False, abund=None, picklist=None):
        _check_select_parameters(ksize=ksize, num=num, abund=abund, moltype=moltype, scaled=scaled)
        if picklist is not None:
            raise NotImplementedError("picklists are outside the GPU path; filter the manifest rows instead")
        rows = self.rows
        if ksize:
            rows = (r for r in rows if r["ksize"] == ksize)
        if moltype:
            rows = (r for r in rows if r["moltype"] == moltype)
        if scaled or containment:
            rows = (r for r in rows if r["scaled"] and not r["num"])
        if num:
            rows = (r for r in rows if r["num"] and not r["scaled"])
        if abund:
            rows = (r for r in rows if r["with_abundance"])
        yield from rows

    def select_to_manifest(self, **kwargs):
        return CollectionManifest(self._select(**kwargs))

    def filter_rows(self, row_filter_fn):
        return CollectionManifest(r for r in self.rows if row_filter_fn(r))

    def filter_on_columns(self, col_filter_fn, col_names):
        return self.filter_rows(lambda row: col_filter_fn([row[c] for c in col_names if row[c] is not None]))

    def locations(self):
        seen = set()
        for row in self.rows:
            loc = row["internal_location"]
            if loc not in seen:
                seen.add(loc)
                yield loc

    def __contains__(self, ss):
        return ss.md5sum() in self._md5_set

    def __add__(self, other):
        mf = CollectionManifest(self.rows)
        mf._add_rows(other.rows)
        return mf

    def __iadd__(self, other):
        if self is other:
            raise Exception("cannot directly add manifest to itself")
        self._add_rows(other.rows)
        return self

    def __bool__(self):
        return bool(self.rows)

    def __len__(self):
        return len(self.rows)

    def __eq__(self, other):
        for a, b in itertools.zip_longest(self.rows, other.rows):
            if a is None or b is None:
                return False
            if any(a[k] != b[k] for k in self.required_keys):
                return False
        return True
