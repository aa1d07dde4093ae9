"""Host emulation of the single-warp CUDA QLFC coders (libbsc_b200/csrc/qlfc_decoder6.cuh, qlfc_fast.cuh, qlfc_adaptive.cuh compiled
with QD3_HOST by tools/qdec3_host.cpp: the same source, the 32 lanes run one after the other) against the oracle, on the CPU.
This pins the LANE LOGIC (which lane owns which MTF slot / row bytes, the staged rows, the branch-free range-coder step)
bit-for-bit; the GPU parity tests (tests/test_gpu_parity.py) then only have to confirm that the device executes it the same way."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "qdec3_host.cpp")
LIB = os.path.join(ROOT, "tools", "bin", "libqdec3_host.so")
DEPS = [SRC] + [os.path.join(ROOT, "libbsc_b200", "csrc", f) for f in ("qlfc_lanes.cuh", "qlfc_decoder6.cuh", "qlfc_decoder6_stream.inc", "qlfc_fast.cuh", "qlfc_adaptive.cuh", "qlfc_tables2.inc", "qlfc_coder.cuh", "qlfc_tables.inc")]


def _hostlib():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        subprocess.run(["g++", "-O2", "-Wno-unknown-pragmas", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
    return ctypes.CDLL(LIB)


@pytest.fixture(scope="module")
def qdec3():
    lib = _hostlib()
    lib.qdec6_host_decode.restype = ctypes.c_int
    lib.qdec6_host_decode.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_int]

    def decode(stream, n, mode):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        out = np.full(n + 64, 0xAA, dtype=np.uint8)
        stats = (ctypes.c_uint * 2)()
        r = lib.qdec6_host_decode(stream.ctypes.data, stream.size, out.ctypes.data, n, stats, mode)
        assert np.all(out[n:] == 0xAA), "wrote past the output slice"
        return r, out[:n], (stats[0], stats[1])
    return decode


def inputs(gen, checker):
    rng = np.random.default_rng(5)

    def bwt(a):
        return checker.bwt_encode(a)[1]
    yield "bwt(text 2M)", bwt(gen.text(2, 2 << 20))
    yield "bwt(text 300k)", bwt(gen.text(7, 300000))
    yield "bwt(skew 300k)", bwt(gen.skew(3, 300000))
    yield "st6-like skew raw", gen.skew(3, 200000)                    # high-entropy: escape mode, wide banks
    yield "alpha4", rng.integers(0, 4, 50000, dtype=np.uint8)
    yield "alpha2", rng.integers(0, 2, 65536, dtype=np.uint8)
    yield "allsame", np.full(3000, 65, dtype=np.uint8)
    yield "zeros", np.zeros(1000, dtype=np.uint8)
    yield "period7", np.tile(np.frombuffer(b"abcabcd", dtype=np.uint8), 3000)
    yield "long runs", np.repeat(rng.integers(0, 200, 3000, dtype=np.uint8), rng.integers(1, 3000, 3000))
    yield "huge run", np.concatenate([gen.text(3, 5000), np.full(3 << 20, 7, np.uint8), gen.text(4, 3000)])
    yield "mixed ranks", np.concatenate([rng.integers(0, 256, 40000, dtype=np.uint8), bwt(gen.text(5, 100000)), rng.integers(0, 256, 40000, dtype=np.uint8)])
    yield "tiny", gen.text(1, 40)


def test_host_emulation_matches_oracle(qdec3, gen, checker, port):
    covered = 0
    for name, a in inputs(gen, checker):
        for enc in (checker, port):
            r, s = enc.encode_block(a)
            if r <= 0:
                continue                                              # not compressible: the container stores it raw
            for mode in (4, 3, 2, 1, 0):                              # the product layouts (five / four streams per SM; tables in global memory), three per SM, diet, full
                n, out, stats = qdec3(s, a.size, mode)
                assert n == a.size, (name, mode, n)
                assert np.array_equal(out, a), (name, mode)
            covered += 1
    assert covered >= 16


def test_host_emulation_rejects_oversized_stream(qdec3, gen, checker):
    a = checker.bwt_encode(gen.text(2, 100000))[1]
    r, s = checker.encode_block(a)
    for mode in (1, 0):
        n, _, _ = qdec3(s, a.size - 1, mode)                          # declared length exceeds the slice
        assert n == -6


# ---- fast coder (coder id 3): libbsc_b200/csrc/qlfc_fast.cuh on the host -------------------------------------------------
def _stream_codec(prefix):
    lib = _hostlib()
    vp, cu = ctypes.c_void_p, ctypes.c_uint
    fdec, fenc = getattr(lib, prefix + "_host_decode"), getattr(lib, prefix + "_host_encode")
    fdec.restype = ctypes.c_int
    fdec.argtypes = [vp, cu, vp, cu, vp]
    fenc.restype = ctypes.c_int
    fenc.argtypes = [vp, vp, vp, cu, cu, vp, vp, cu, vp]

    def decode(stream, n):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        out = np.full(n + 64, 0xAA, dtype=np.uint8)
        r = fdec(stream.ctypes.data, stream.size, out.ctypes.data, n, None)
        assert np.all(out[n:] == 0xAA), "wrote past the output slice"
        return r, out[:n]

    def encode(a, ranks, mtf, out_cap=None):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        heads = np.concatenate([[0], np.flatnonzero(np.diff(a) != 0) + 1]).astype(np.uint32)
        run_pos = np.concatenate([heads, [a.size]]).astype(np.uint32)
        run_sym = np.ascontiguousarray(a[heads])
        assert ranks.size == heads.size
        out_cap = a.size if out_cap is None else out_cap
        out = np.full(a.size + 4096 + 256, 0xAA, dtype=np.uint8)
        mtf = np.ascontiguousarray(mtf, dtype=np.uint8); ranks = np.ascontiguousarray(ranks, dtype=np.uint8)
        r = fenc(run_pos.ctypes.data, run_sym.ctypes.data, ranks.ctypes.data, heads.size, a.size, mtf.ctypes.data, out.ctypes.data, out_cap, None)
        return r, (out[:r].copy() if r > 0 else None)
    return decode, encode


@pytest.fixture(scope="module")
def qfast():
    return _stream_codec("qfast")


def _check_stream_codec(codec, coder, gen, checker, port):
    decode, encode = codec
    covered = 0
    for name, a in list(inputs(gen, checker)) + [("rand", gen.rand(1, 70000))]:
        r_ref, s_ref = checker.encode_block(a, coder=coder)
        ranks, mtf = port.transform(a)
        r, s = encode(a, ranks, mtf)
        assert r == r_ref, (name, r, r_ref)                            # same length, or the same NOT_COMPRESSIBLE (-3)
        if r_ref > 0:
            assert np.array_equal(s, s_ref), name
            n, out = decode(s_ref, a.size)
            assert n == a.size and np.array_equal(out, a), name
            covered += 1
    assert covered >= 10
    a = checker.bwt_encode(gen.text(2, 100000))[1]
    assert decode(checker.encode_block(a, coder=coder)[1], a.size - 1)[0] == -6


def test_fast_coder_host_emulation_matches_oracle(qfast, gen, checker, port):
    _check_stream_codec(qfast, 3, gen, checker, port)


# ---- adaptive coder (coder id 2): libbsc_b200/csrc/qlfc_adaptive.cuh on the host ---------------------------------------
def test_adaptive_coder_host_emulation_matches_oracle(gen, checker, port):
    lib = _hostlib()
    lib.qadapt_smem_bytes.restype = ctypes.c_uint
    assert lib.qadapt_smem_bytes() <= 227 * 1024
    _check_stream_codec(_stream_codec("qadapt"), 2, gen, checker, port)


# ---- layout-templated decoder (qlfc_decoder6.cuh): full layout = refactoring check, diet layout = two streams per SM -------------
def test_layout_templated_decoder_host_emulation(gen, checker, port):
    lib = _hostlib()
    lib.qdec6_host_decode.restype = ctypes.c_int
    lib.qdec6_host_decode.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_int]
    lib.qdec6_smem_bytes.restype = ctypes.c_uint
    assert lib.qdec6_smem_bytes(1) <= 113 * 1024 < lib.qdec6_smem_bytes(0)       # diet: two CTAs per SM; full: one
    for layout, per_sm in ((2, 3), (3, 4), (4, 5)):                              # 228 KB per SM, 1 KB reserved per CTA
        assert lib.qdec6_smem_bytes(layout) + 1024 <= 228 * 1024 // per_sm, (layout, per_sm)
    covered, rare = 0, [0, 0]
    for name, a in inputs(gen, checker):
        r, s = checker.encode_block(a)
        if r <= 0:
            continue
        for layout in (0, 1, 2, 3, 4):
            out = np.full(a.size + 64, 0xAA, dtype=np.uint8)
            stats = (ctypes.c_uint * 2)()
            s_ = np.ascontiguousarray(s)
            n = lib.qdec6_host_decode(s_.ctypes.data, s_.size, out.ctypes.data, a.size, stats, layout)
            assert n == a.size and np.array_equal(out[:a.size], a) and np.all(out[a.size:] == 0xAA), (name, layout)
            if layout < 2:
                rare[layout] += stats[0]
        covered += 1
    assert covered >= 8
    assert rare[1] > rare[0]                                                      # the diet layout really has more row events (stats[0] = 2 per event)


def test_escape_mode_rows_instead_of_cache_misses(gen, checker):
    """High-entropy streams (avgRank >= 32) code every rank bit in the escape bank: 2 x 64 K counters per stream.  q_decode6 fetches
    the two 512-byte rows a run needs (state row, symbol row) whole instead of sending 16 accesses per run through write-back
    caches (7-10 misses per run, measured with this emulation before the change); it uses rows for every non-resident bank.  Checked here: bit-exact, exactly two row
    fetches per escape-mode run, and (almost) no cache miss on such a stream."""
    lib = _hostlib()
    lib.qdec6_host_decode.restype = ctypes.c_int
    lib.qdec6_host_decode.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_int]
    a = gen.skew(3, 600000)
    r, s = checker.encode_block(a)
    assert r > 0
    s_ = np.ascontiguousarray(s)
    runs = int((np.diff(a.astype(np.int16)) != 0).sum() + 1)
    for layout in (0, 1, 2):
        out = np.full(a.size + 64, 0xAA, dtype=np.uint8)
        stats = (ctypes.c_uint * 2)()
        n = lib.qdec6_host_decode(s_.ctypes.data, s_.size, out.ctypes.data, a.size, stats, layout)
        assert n == a.size and np.array_equal(out[:a.size], a) and np.all(out[a.size:] == 0xAA), layout
        # (the first runs, before avgRank has climbed to 32, still take the plain path and a few cached exponents)
        assert 1.9 * runs <= stats[0] <= 2 * runs + 200 and stats[1] < 200, (layout, stats[0], stats[1], runs)
