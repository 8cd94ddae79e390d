#!/usr/bin/env python
"""Writes tests/golden/oracle_hashes.json: sha256 of the oracle's decoded pixels (scalar and SSSE3 arithmetic) for every
fixture, plus the libjpeg-turbo (PIL) decode distance.  The golden PNGs of the reference pin the oracle to +-3 (they are
libjpeg's pixels, tests/reftest/mod.rs:99); these hashes freeze the exact bytes the oracle produced when it was
validated, so that a later edit of oracle/*.c cannot drift inside that tolerance unnoticed.
Re-run only when the oracle is deliberately changed, and say why in the commit."""
import glob
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

out = {}
files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "reftest", "**", "*.jpg"), recursive=True))
files += sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "benches", "*.jpg")))
for p in files:
    rel = os.path.relpath(p, os.path.join(ROOT, "tests", "golden"))
    data = open(p, "rb").read()
    entry = {}
    for name, a in (("scalar", oracle.ARITH_SCALAR), ("ssse3", oracle.ARITH_SSSE3)):
        try:
            px = oracle.Decoder(data, a).decode()
            entry[name] = {"sha256": hashlib.sha256(px.tobytes()).hexdigest(), "bytes": int(px.size)}
        except oracle.OracleError as e:
            entry[name] = {"error": int(e.code)}
    out[rel] = entry
path = os.path.join(ROOT, "tests", "golden", "oracle_hashes.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print("wrote %s: %d files" % (path, len(out)))
