"""CPU-only tests: the oracle is pinned against the known-answer values of SURVEY.md Appendix C and
against the unmodified reference (oracle/_ref, when built); the C-ABI library loads and exports
every symbol include/libbsc_b200.h declares.  No GPU work here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- generators (Appendix C adler32 values) -------------------------------------------------
def test_generators_match_appendix_c(gen):
    assert gen.adler32(gen.rand(1, 1048576)) == 0x90fa0fec
    assert gen.adler32(gen.skew(3, 1 << 22)) == gen.adler32(gen.skew(3, 33554432)[: 1 << 22])
    t = gen.text(2, 4 << 20)
    assert bytes(t[:8]) == b"Oumrefue"
    assert gen.adler32(t) == gen.adler32(gen.text(2, 6 << 20)[: 4 << 20])     # sequential generator


@pytest.mark.slow
def test_generators_full_kats(gen):
    assert gen.adler32(gen.skew(3, 33554432)) == 0x82fefea1
    t = gen.text(2, 67108864)
    assert gen.adler32(t) == 0x2762809f
    assert gen.adler32(t[:26214400]) == 0x5cf8c0f7


# ---- K1 known answers (1 MiB random) against the port ---------------------------------------
def test_port_k1(gen, port):
    a = gen.rand(1, 1048576)
    idx, L, aux = port.bwt_encode(a)
    assert idx == 791385
    assert aux == [633903, 252159, 118572, 805028, 426935, 565037, 575571]
    assert gen.adler32(L) == 0x5ce80fec
    r, _ = port.coder_compress(L, 1, 3)
    assert r == -3
    r, blk = port.compress(a, 1, 1, 3)
    assert r == 1048604 and gen.adler32(blk) == 0x242115a8
    assert int.from_bytes(bytes(blk[8:12]), "little") == 0                   # stored block


def test_port_k5_prefix_st6(gen, port, ref):
    a = gen.skew(3, 1 << 20)
    i1, L1 = port.st_encode(a, 6)
    i2, L2 = ref.st_encode(a, 6)
    assert i1 == i2 and np.array_equal(L1, L2)


# ---- port vs unmodified reference -----------------------------------------------------------
def _inputs(gen):
    rng = np.random.default_rng(0)
    yield "text300k", gen.text(7, 300000)
    yield "text1M", gen.text(2, 1 << 20)
    yield "skew300k", gen.skew(3, 300000)
    yield "rand70k", gen.rand(1, 70000)
    yield "alpha4", rng.integers(0, 4, 5000, dtype=np.uint8)
    yield "allsame", np.full(3000, 65, dtype=np.uint8)
    yield "zeros", np.zeros(1000, dtype=np.uint8)
    yield "tiny29", gen.text(1, 29)
    yield "tiny100", gen.text(1, 100)
    yield "withzeros", np.concatenate([gen.text(3, 5000), np.zeros(40, np.uint8), gen.text(4, 3000), np.zeros(9, np.uint8)])


def test_port_matches_reference_stages(gen, port, ref):
    for name, a in _inputs(gen):
        r1, L1, i1 = port.bwt_encode(a)
        r2, L2, i2 = ref.bwt_encode(a)
        assert r1 == r2 and i1 == i2 and np.array_equal(L1, L2), name
        d, T = port.bwt_decode(L2, r2)
        assert d == 0 and np.array_equal(T, a), name
        for feats in (1, 3):
            c1, s1 = port.coder_compress(L2, 1, feats)
            c2, s2 = ref.coder_compress(L2, 1, feats)
            assert c1 == c2, (name, feats)
            if c2 > 0:
                assert np.array_equal(s1, s2), (name, feats)
                n1, o1 = port.coder_decompress(s2, a.size)
                assert n1 == a.size and np.array_equal(o1, L2), name
        for k in (3, 4, 5, 6):
            x1, y1 = port.st_encode(a, k)
            x2, y2 = ref.st_encode(a, k)
            assert x1 == x2 and np.array_equal(y1, y2), (name, k)
            d1, t1 = port.st_decode(y2, k, x2)            # inverse ST: port and reference both give the input back
            d2, t2 = ref.st_decode(y2, k, x2)
            assert d1 == 0 and d2 == 0 and np.array_equal(t1, a) and np.array_equal(t2, a), (name, k)
        for k in (7, 8):                                  # forward k = 7, 8 exists only in the reference's CUDA build; its inverse is on the CPU
            x1, y1 = port.st_encode(a, k)
            d2, t2 = ref.st_decode(y1, k, x1)
            d1, t1 = port.st_decode(y1, k, x1)
            assert d1 == 0 and d2 == 0 and np.array_equal(t1, a) and np.array_equal(t2, a), (name, k)


def test_port_matches_reference_blocks(gen, port, ref):
    for name, a in _inputs(gen):
        for sorter in (1, 5):
            z1, b1 = port.compress(a, sorter, 1, 3)
            z2, b2 = ref.compress(a, sorter, 1, 3)
            assert z1 == z2 and np.array_equal(b1, b2), (name, sorter)
        for sorter in (1, 4, 6):
            z, b = ref.compress(a, sorter, 1, 3)
            q, u = port.decompress(b)
            assert q == 0 and np.array_equal(u, a), (name, sorter)


def test_port_transform_and_split(gen, port, ref):
    a = gen.text(5, 400000)
    _, L, _ = ref.bwt_encode(a)
    ranks, mtf = port.transform(L)
    assert ranks.size > 0 and ranks[-1] == 1
    st, sz = port.split_blocks(L, 2)
    assert st[0] == 0 and st[1] == sz[0] and sz[0] + sz[1] == L.size
    # single sub-block streams of the port equal the reference's
    for piece in (L[:st[1]], L[st[1]:]):
        r1, s1 = port.encode_block(piece)
        r2, s2 = ref.encode_block(piece)
        assert r1 == r2 and np.array_equal(s1, s2)


@pytest.mark.parametrize("coder", [2, 3])
def test_port_other_coders_match_reference(gen, port, ref, coder):
    """Coder ids 2 (adaptive `-e2`, qlfc.cpp:463-823 / 1366-1666, logistic mixing predictor.h:74-213) and 3 (fast `-e0`,
    qlfc.cpp:1135-1336 / 1933-2127): streams, container and blocks equal the reference's -- for the fast coder including
    the header's unnormalised-bit quirk (rangecoder.h:165-177 called with `c & (1 << bit)`)."""
    rng = np.random.default_rng(5)
    extra = [("long runs", np.repeat(rng.integers(0, 200, 2000, dtype=np.uint8), rng.integers(1, 3000, 2000))),
             ("huge run", np.concatenate([gen.text(3, 5000), np.full(3 << 20, 7, np.uint8), gen.text(4, 3000)])),
             ("alpha2", rng.integers(0, 2, 65536, dtype=np.uint8))]
    for name, a in list(_inputs(gen)) + extra:
        L = ref.bwt_encode(a)[1] if a.size < (1 << 20) + 1 else a
        r1, s1 = port.encode_block(L, coder=coder)
        r2, s2 = ref.encode_block(L, coder=coder)
        assert r1 == r2, (name, r1, r2)
        if r2 > 0:
            assert np.array_equal(s1, s2), name
            n, o = port.decode_block(s2, L.size, coder=coder)
            assert n == L.size and np.array_equal(o, L), name
        for feats in (1, 3):
            c1, t1 = port.coder_compress(L, coder, feats)
            c2, t2 = ref.coder_compress(L, coder, feats)
            assert c1 == c2, (name, feats)
            if c2 > 0:
                assert np.array_equal(t1, t2), (name, feats)
                n, o = port.coder_decompress(t2, L.size, coder)
                assert n == L.size and np.array_equal(o, L), name
    a = gen.text(2, 3 << 20)
    z1, b1 = port.compress(a, 1, coder, 3)
    z2, b2 = ref.compress(a, 1, coder, 3)
    assert z1 == z2 and np.array_equal(b1, b2)
    assert port.decompress(b2)[0] == 0 and ref.decompress(b1)[0] == 0


def test_port_undoes_reference_lzp(gen, port, ref):
    """The LZP stage stays on the host (BASELINE.json north_star); only its DECODER is restated (lzp.cpp:564-674, 813-887), so
    that blocks made with the reference's default options (`lzpHashSize` 15, `lzpMinLen` 128) can be decoded."""
    rng = np.random.default_rng(1)
    rep = np.tile(gen.text(3, 700), 900)                                          # long repeats: many LZP matches
    mixed = np.concatenate([gen.text(4, 300000), rep[:400000], np.full(5000, 0xF2, np.uint8), gen.text(4, 300000), rng.integers(0, 256, 50000, dtype=np.uint8)])
    for name, a in (("rep", rep), ("mixed", mixed), ("text5M", np.tile(gen.text(2, 1 << 20), 5)), ("flags", np.full(100000, 0xF2, np.uint8))):
        for h, m in ((15, 128), (12, 4), (16, 32), (18, 255)):
            r, s = ref.lzp_compress(a, h, m)
            if r <= 0:
                continue
            n, out = port.lzp_decompress(s, a.size, h, m)
            assert n == a.size and np.array_equal(out, a), (name, h, m)
            assert port.lzp_decompress(s, a.size - 1, h, m)[0] < 0                 # never writes past the capacity
        z, blk = ref.compress_lzp(a)                                              # the reference's default block
        assert z > 0 and (int.from_bytes(bytes(blk[8:12]), "little") >> 8) != 0, name
        q, u = port.decompress(blk)
        assert q == 0 and np.array_equal(u, a), name


def test_product_host_lzp_decoder_matches_reference(gen, ref):
    """libbsc_b200's host-side inverse LZP stage (csrc/lzp_host.h, what bsc_decompress runs after the GPU stages) against the
    reference's bsc_lzp_compress -- host-only code, so it is checked here without a GPU."""
    import libbsc_b200
    L = libbsc_b200.lib()
    rng = np.random.default_rng(2)
    rep = np.tile(gen.text(3, 700), 900)
    mixed = np.concatenate([gen.text(4, 300000), rep[:400000], np.full(5000, 0xF2, np.uint8), rng.integers(0, 256, 50000, dtype=np.uint8)])
    for name, a in (("rep", rep), ("mixed", mixed), ("text5M", np.tile(gen.text(2, 1 << 20), 5))):
        for h, m in ((15, 128), (12, 4), (18, 255)):
            r, s = ref.lzp_compress(a, h, m)
            if r <= 0:
                continue
            out = np.full(a.size + 64, 0xAA, dtype=np.uint8)
            n = L.bscb200_lzp_decompress_host(s.ctypes.data, s.size, out.ctypes.data, a.size, h, m)
            assert n == a.size and np.array_equal(out[:a.size], a) and np.all(out[a.size:] == 0xAA), (name, h, m)
            assert L.bscb200_lzp_decompress_host(s.ctypes.data, s.size, out.ctypes.data, a.size - 1, h, m) == -6
            bad = s.copy(); bad[bad.size // 2] ^= 0x55
            out2 = np.full(a.size + 64, 0xAA, dtype=np.uint8)
            L.bscb200_lzp_decompress_host(bad.ctypes.data, bad.size, out2.ctypes.data, a.size, h, m)   # any result, but inside the capacity
            assert np.all(out2[a.size:] == 0xAA)
    assert L.bscb200_lzp_decompress_host(s.ctypes.data, s.size, out.ctypes.data, a.size, 9, 128) == -1


def test_product_host_lzp_encoder_matches_reference(gen, ref):
    """The forward LZP stage of libbsc_b200 (csrc/lzp_host.h: the all five variants an x86-64 reference build picks
    from by (hashSize, minLen) -- small, small2x, medium, large, generic -- and both chunk container rules) reproduces
    bsc_lzp_compress byte for byte."""
    import libbsc_b200
    L = libbsc_b200.lib()
    rng = np.random.default_rng(2)
    rep = np.tile(gen.text(3, 700), 900)
    cases = {"rep": rep, "mixed": np.concatenate([gen.text(4, 300000), rep[:400000], np.full(5000, 0xF2, np.uint8), rng.integers(0, 256, 50000, dtype=np.uint8)]),
             "text5M": np.tile(gen.text(2, 1 << 20), 5), "flags": np.full(100000, 0xF2, np.uint8), "small": np.tile(gen.text(5, 300), 100), "rand": gen.rand(1, 500000),
             "nearrep": np.concatenate([gen.text(9, 200000)] * 3 + [gen.rand(2, 100000)] + [gen.text(9, 200000)] * 2)}
    for name, a in cases.items():
        for h, m in ((15, 128), (16, 64), (17, 17), (12, 255), (20, 128), (18, 4),
                     (15, 4), (13, 8), (17, 16), (15, 6), (10, 12), (16, 9), (15, 15)):
            for feats in (1, 3):
                r1, s1 = ref.lzp_compress(a, h, m, feats)
                out = np.full(a.size + 64, 0xAA, dtype=np.uint8)
                r2 = L.bscb200_lzp_compress_host(a.ctypes.data, out.ctypes.data, a.size, h, m, feats)
                assert r1 == r2, (name, h, m, feats, r1, r2)
                assert r1 <= 0 or np.array_equal(out[:r2], s1), (name, h, m, feats)
                assert np.all(out[a.size:] == 0xAA), "wrote past n bytes"
    a = cases["rep"]
    out = np.empty(a.size + 64, dtype=np.uint8)
    assert L.bscb200_lzp_compress_host(a.ctypes.data, out.ctypes.data, a.size, 9, 128, 3) == -1


def test_small_n_bwt_conventions(port, ref):
    for n in (0, 1, 2, 3, 7, 8, 15, 16, 17, 31, 33):
        a = (np.arange(n, dtype=np.uint8) * 7 + 3) % 5
        assert port.bwt_encode(a, aux=True)[0] == ref.bwt_encode(a, aux=True)[0], n
        r1, L1, _ = port.bwt_encode(a, aux=False)
        r2, L2, _ = ref.bwt_encode(a, aux=False)
        assert r1 == r2 and np.array_equal(L1, L2), n


# ---- boundary: the library loads and exports the declared ABI -------------------------------
def test_library_exports_declared_abi():
    import libbsc_b200
    assert os.path.exists(libbsc_b200.LIB_PATH), "run __graft_entry__.build() first"
    L = ctypes.CDLL(libbsc_b200.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "libbsc_b200.h")).read()
    declared = set(re.findall(r"\b(bsc_\w+|bscb200_\w+)\s*\(", header))
    assert len(declared) >= 35
    for name in sorted(declared):
        assert hasattr(L, name), "missing export: " + name
    assert declared == set(libbsc_b200.EXPORTED_SYMBOLS)


def test_host_only_entry_points_need_no_gpu():
    """Pure header logic (no device work): bsc_block_info validation mirrors libbsc.cpp:340-418."""
    import libbsc_b200
    L = libbsc_b200.lib()
    hdr = np.zeros(28, dtype=np.uint8)
    assert L.bsc_block_info(hdr.ctypes.data, 10, None, None, 0) == -5          # UNEXPECTED_EOB
    assert L.bsc_block_info(hdr.ctypes.data, 28, None, None, 0) == -6          # bad header adler


# ---- seeded sweep: many small inputs of random size / alphabet / structure, port vs reference, every stage -------------------
def _sweep_inputs(count, seed):
    rng = np.random.default_rng(seed)
    for i in range(count):
        n = int(rng.integers(1, 6000)) if i % 7 else int(rng.integers(1, 40))
        sigma = int(rng.choice([1, 2, 3, 4, 16, 64, 256]))
        kind = i % 5
        if kind == 0:                                       # i.i.d. symbols
            a = rng.integers(0, sigma, n, dtype=np.uint8)
        elif kind == 1:                                     # long runs
            a = np.repeat(rng.integers(0, sigma, n, dtype=np.uint8), rng.integers(1, 40, n))[:n].astype(np.uint8)
        elif kind == 2:                                     # periodic with a defect
            p = rng.integers(0, sigma, int(rng.integers(1, 12)), dtype=np.uint8)
            a = np.tile(p, n // p.size + 1)[:n].copy(); a[int(rng.integers(0, n))] ^= 1
        elif kind == 3:                                     # repeats of earlier material (LZ-like)
            a = rng.integers(0, sigma, n, dtype=np.uint8)
            for _ in range(4):
                if n > 8:
                    s, d, l = (int(x) for x in (rng.integers(0, n - 4), rng.integers(0, n - 4), rng.integers(1, n // 2)))
                    l = min(l, n - s, n - d); a[d:d + l] = a[s:s + l].copy()
        else:                                               # high bytes, zeros mixed in
            a = rng.integers(200, 256, n, dtype=np.uint8); a[rng.integers(0, n, max(1, n // 9))] = 0
        yield "sweep%d(n=%d,sigma=%d,kind=%d)" % (i, n, sigma, kind), np.ascontiguousarray(a, dtype=np.uint8)


def test_port_matches_reference_on_a_seeded_sweep(port, ref):
    for name, a in _sweep_inputs(160, 20260924):
        aux = a.size >= 16                                   # secondary indexes need n >= 16 (test_small_n_bwt_conventions)
        r1, L1, i1 = port.bwt_encode(a, aux)
        r2, L2, i2 = ref.bwt_encode(a, aux)
        assert r1 == r2 and i1 == i2 and np.array_equal(L1, L2), name
        if a.size > 1:
            d, T = port.bwt_decode(L2, r2)
            assert d == 0 and np.array_equal(T, a), name
        for coder in (1, 2, 3):
            c1, s1 = port.coder_compress(L2, coder, 3)
            c2, s2 = ref.coder_compress(L2, coder, 3)
            assert c1 == c2, (name, coder)
            if c2 > 0:
                assert np.array_equal(s1, s2), (name, coder)
                n1, o1 = port.coder_decompress(s2, a.size, coder)
                assert n1 == a.size and np.array_equal(o1, L2), (name, coder)
        k = 3 + (a.size % 4)
        x1, y1 = port.st_encode(a, k)
        x2, y2 = ref.st_encode(a, k)
        assert x1 == x2 and np.array_equal(y1, y2), (name, k)
        if x2 >= 0:
            d1, t1 = port.st_decode(y2, k, x2)
            assert d1 == 0 and np.array_equal(t1, a), (name, k)
        sorter = (1, 3, 4, 5, 6)[a.size % 5]
        z1, b1 = port.compress(a, sorter, 1, 3)
        z2, b2 = ref.compress(a, sorter, 1, 3)
        assert z1 == z2 and np.array_equal(b1, b2), (name, sorter)
        q, u = port.decompress(b2)
        assert q == 0 and np.array_equal(u, a), (name, sorter)
