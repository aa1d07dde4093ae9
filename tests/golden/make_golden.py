#!/usr/bin/env python3
"""Generates tests/golden/*.bin from the UNMODIFIED reference compiled here (oracle/_ref/libbsc_ref.so, built by
oracle/Makefile from /root/reference).  Run in the build container only; the fixtures are committed so that the parity
tests have reference outputs on machines that have neither /root/reference nor oracle/_ref.

Inputs are the seeded generators of SURVEY.md Appendix C (tools/bscgen.c), 20 000 bytes each; every fixture file is the
raw output of one reference call:
    <input>.bwt            int32 LE primary index, then L                       (bsc_bwt_encode)
    <input>.st<k>          int32 LE index, then L, k = 3..6                     (bsc_st_encode; k = 7, 8 need the CUDA build)
    <input>.coder<c>       the container of bsc_coder_compress(L_bwt, coder c)  c = 1 static, 2 adaptive, 3 fast
    <input>.block.m<s>e<c> the block of bsc_compress(sorter s, coder c)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle

N = 20000


def inputs(gen):
    return {"text7": gen.text(7, N), "skew3": gen.skew(3, N), "rand1": gen.rand(1, N)}


def main():
    gen, ref = pyoracle.Gen(), pyoracle.Ref()
    for name, a in inputs(gen).items():
        def put(suffix, *parts):
            with open(os.path.join(HERE, name + "." + suffix), "wb") as f:
                for p in parts:
                    f.write(p if isinstance(p, bytes) else np.ascontiguousarray(p).tobytes())
        r, L, _ = ref.bwt_encode(a)
        put("bwt", np.int32(r).tobytes(), L)
        for k in (3, 4, 5, 6):
            i, Ls = ref.st_encode(a, k)
            put("st%d" % k, np.int32(i).tobytes(), Ls)
        for c in (1, 2, 3):
            z, s = ref.coder_compress(L, c, 3)
            put("coder%d" % c, s if z > 0 else np.int32(z).tobytes())
            for sorter in (1, 6):
                z, b = ref.compress(a, sorter, c, 3)
                put("block.m%de%d" % (sorter, c), b)
    print("ok")


if __name__ == "__main__":
    main()
