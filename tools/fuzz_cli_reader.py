#!/usr/bin/env python3
"""tools/fuzz_cli_reader.py -- the archive reader of libbsc_b200/cli/bsc_b200.cpp under ASan/UBSan (CPU only).

The CLI source is built against the unmodified reference library (-DBSCB200_CLI_REF) with -fsanitize=address,undefined and fed
archives of the reference CLI (plain, LZP, reordered/reversed, segmented) with flipped bytes, truncations and overwritten size
fields.  Every mutated archive must be rejected (or decoded) without a sanitizer report or a signal.  Last run: 240 cases, clean.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402

D = "/tmp/bscb200_clifuzz"
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def main():
    os.makedirs(D, exist_ok=True)
    exe = os.path.join(D, "cli_asan")
    subprocess.run(["/usr/bin/g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-std=c++17", "-pthread", "-fopenmp",
                    "-DBSCB200_CLI_REF", os.path.join(ROOT, "libbsc_b200", "cli", "bsc_b200.cpp"), "-o", exe, "-L" + REFDIR, "-lbsc_ref",
                    "-Wl,-rpath," + REFDIR], check=True)
    gen = pyoracle.Gen()
    src = os.path.join(D, "in.bin")
    np.concatenate([gen.text(3, 700000), gen.rand(1, 50000), gen.text(4, 400000)]).tofile(src)
    archives = []
    for i, opts in enumerate((["-b1", "-p"], ["-b1"], ["-b1", "-p", "-r", "-cp"], ["-b1", "-s", "-ca"])):
        p = os.path.join(D, "a%d.bsc" % i)
        subprocess.run([os.path.join(REFDIR, "bsc"), "e", src, p, "-t"] + opts, check=True, capture_output=True)
        archives.append(p)
    rng = np.random.default_rng(5)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    cases = bad = 0
    for p in archives:
        raw = np.fromfile(p, dtype=np.uint8)
        r = subprocess.run([exe, "d", p, os.path.join(D, "ok.out"), "-j2"], capture_output=True, env=env)
        assert r.returncode == 0 and open(os.path.join(D, "ok.out"), "rb").read() == open(src, "rb").read(), "unmutated archive must decode"
        for t in range(60):
            m = raw.copy()
            kind = t % 4
            if kind == 0:                                   # container header / first block headers
                for _ in range(3):
                    m[rng.integers(0, min(200, m.size))] ^= rng.integers(1, 256)
            elif kind == 1:
                m = m[:rng.integers(0, m.size)]
            elif kind == 2:
                for _ in range(5):
                    m[rng.integers(0, m.size)] ^= rng.integers(1, 256)
            else:                                           # size / offset fields
                o = rng.integers(0, max(1, m.size - 4))
                m[o:o + 4] = np.frombuffer(np.uint32(rng.choice([0xffffffff, 0x7fffffff, 0x80000000, 1, 0])).tobytes(), np.uint8)
            mp = os.path.join(D, "m.bsc")
            m.tofile(mp)
            r = subprocess.run([exe, "d", mp, os.path.join(D, "m.out"), "-j2"], capture_output=True, text=True, env=env, timeout=300)
            cases += 1
            if "AddressSanitizer" in r.stderr or "runtime error" in r.stderr or r.returncode < 0:
                bad += 1
                print("FAIL", p, t, kind, r.returncode, r.stderr[-500:])
    print("cases", cases, "failures", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
