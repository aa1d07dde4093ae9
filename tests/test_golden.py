"""Golden vectors: tests/golden/*.bin are raw outputs of the UNMODIFIED reference (generated here by
tests/golden/make_golden.py from oracle/_ref; the reference ships no fixtures of its own).
They pin the oracle port on machines without the reference, and the CUDA path on the GPU box."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N = 20000


def _inputs(gen):
    return {"text7": gen.text(7, N), "skew3": gen.skew(3, N), "rand1": gen.rand(1, N)}


def _load(name, suffix):
    return np.fromfile(os.path.join(GOLD, name + "." + suffix), dtype=np.uint8)


def _index_and_bytes(raw):
    return int(np.frombuffer(raw[:4].tobytes(), dtype="<i4")[0]), raw[4:]


def _check_impl(impl, gen, coders, st_ks, sorters):
    for name, a in _inputs(gen).items():
        idx, L = _index_and_bytes(_load(name, "bwt"))
        r, L1, _ = impl.bwt_encode(a)
        assert r == idx and np.array_equal(L1, L), name
        d, T = impl.bwt_decode(L, idx)
        assert d == 0 and np.array_equal(T, a), name
        for k in st_ks:
            i, Ls = _index_and_bytes(_load(name, "st%d" % k))
            i1, L1 = impl.st_encode(a, k)
            assert i1 == i and np.array_equal(L1, Ls), (name, k)
            d, T = impl.st_decode(Ls, k, i)
            assert d == 0 and np.array_equal(T, a), (name, k)
        for c in coders:
            gold = _load(name, "coder%d" % c)
            z, s = impl.coder_compress(L, c, 3)
            if gold.size == 4 and int(np.frombuffer(gold.tobytes(), dtype="<i4")[0]) < 0:
                assert z == int(np.frombuffer(gold.tobytes(), dtype="<i4")[0]), (name, c)     # not compressible
            else:
                assert z == gold.size and np.array_equal(s, gold), (name, c)
                n, out = impl.coder_decompress(gold, L.size, c)
                assert n == L.size and np.array_equal(out, L), (name, c)
            for sorter in sorters:
                blk = _load(name, "block.m%de%d" % (sorter, c))
                z, b = impl.compress(a, sorter, c, 3)
                assert z == blk.size and np.array_equal(b[:z], blk), (name, sorter, c)
                q, u = impl.decompress(blk)
                assert q == 0 and np.array_equal(u, a), (name, sorter, c)


def test_fixtures_present():
    assert len([f for f in os.listdir(GOLD) if f.endswith(tuple("0123456789t"))]) >= 40


def test_oracle_port_reproduces_reference_fixtures(gen, port):
    _check_impl(port, gen, coders=(1, 2, 3), st_ks=(3, 4, 5, 6), sorters=(1, 6))


@pytest.mark.gpu
def test_cuda_path_reproduces_reference_fixtures(gen, bsc):
    _check_impl(bsc, gen, coders=(1, 2, 3), st_ks=(3, 4, 5, 6), sorters=(1, 6))
