#!/usr/bin/env python3
"""tools/sanity_small.py -- every stage on a few SMALL inputs, checked against the oracle.
Meant to be run under compute-sanitizer (memcheck / racecheck / initcheck), where the full GPU test
suite would take too long:  compute-sanitizer --tool memcheck python tools/sanity_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libbsc_b200
from oracle import pyoracle

gen = pyoracle.Gen()
chk = pyoracle.best()
port = pyoracle.Port()
bsc = libbsc_b200.Bsc(3)
rng = np.random.default_rng(7)
cases = [("text300k", gen.text(7, 300000)), ("skew40k", gen.skew(3, 40000)), ("alpha4", rng.integers(0, 4, 5000, dtype=np.uint8)),
         ("allsame", np.full(3000, 65, dtype=np.uint8)), ("tiny100", gen.text(1, 100)), ("ragged4097", gen.text(8, 4097))]
bad = 0
for name, a in cases:
    r2, L2, x2 = chk.bwt_encode(a); r1, L1, x1 = bsc.bwt_encode(a)
    ok = r1 == r2 and x1 == x2 and np.array_equal(L1, L2)
    d, T = bsc.bwt_decode(L2, r2); ok &= d == 0 and np.array_equal(T, a)
    for k in (3, 6, 8):
        i2, S2 = (chk if k <= 6 else port).st_encode(a, k); i1, S1 = bsc.st_encode(a, k)
        ok &= i1 == i2 and np.array_equal(S1, S2)
    for feats in (3, 1):
        c2, s2 = chk.coder_compress(L2, 1, feats); c1, s1 = bsc.coder_compress(L2, 1, feats)
        ok &= c1 == c2 and (c2 <= 0 or np.array_equal(s1, s2))
    if c2 > 0:
        nn, out = bsc.coder_decompress(s2, L2.size); ok &= nn == L2.size and np.array_equal(out, L2)
    z2, b2 = chk.compress(a, 1, 1, 3); z1, b1 = bsc.compress(a, 1, 1, 3)
    ok &= z1 == z2 and np.array_equal(b1, b2)
    q, u = bsc.decompress(b2); ok &= q == 0 and np.array_equal(u, a)
    print(name, "ok" if ok else "MISMATCH", flush=True)
    bad += not ok
print("sanity_small:", "all ok" if not bad else "%d failures" % bad)
sys.exit(1 if bad else 0)
