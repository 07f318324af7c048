#!/usr/bin/env python
"""Record an ncu-measured constant in profiles/ncu_traffic.json together with the git blob ids of the kernel
source files it was measured on (bench.py reports it as roofline.traffic only while those files are unchanged).

    python scripts/record_traffic.py <key> <value> "<source: profile file, what was summed>" <file> [<file> ...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import git_blob_hash  # noqa: E402

key, value, source, files = sys.argv[1], float(sys.argv[2]), sys.argv[3], sys.argv[4:]
assert files, "name the kernel source files the number depends on"
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
with open(path) as fh:
    doc = json.load(fh)
doc.setdefault("entries", {})[key] = {"value": value, "source": source,
                                      "files": {f: git_blob_hash(os.path.join(ROOT, f)) for f in files}}
with open(path, "w") as fh:
    json.dump(doc, fh, indent=1)
print(key, doc["entries"][key])
