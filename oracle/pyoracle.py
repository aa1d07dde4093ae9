"""oracle/pyoracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes bindings for
  * liboracle.so          our plain-C restatement (oracle/bsc_oracle.c), class `Port`
  * _ref/libbsc_ref.so    the UNMODIFIED reference compiled from /root/reference by
                          oracle/Makefile, class `Ref` (present when it was built in the build
                          container; travels to the GPU box as a built file)
and for tools/libbscgen.so (the deterministic input generators of SURVEY.md Appendix C).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (libbsc_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
vp, ci = ctypes.c_void_p, ctypes.c_int

os.environ.setdefault("OMP_WAIT_POLICY", "passive")   # SURVEY.md section 6 warning


def _ptr(a):
    return a.ctypes.data if a is not None else None


def build(force=False):
    """(Re)build liboracle.so, libbscgen.so and -- when /root/reference exists -- _ref/libbsc_ref.so."""
    # Through make wherever the sources can change (the build container: /root/reference present), so that staleness is handled
    # there.  On the GPU box the prebuilt files of the snapshot are used as they are: nothing is rebuilt on the strength of file
    # times a copy may not have preserved, and the ranks of a torchrun launch never race to rewrite a library another rank loads.
    have = os.path.exists(os.path.join(HERE, "liboracle.so"))
    if force or not have or os.path.exists("/root/reference/libbsc/libbsc.h"):
        try:
            subprocess.check_call(["make", "-C", HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
        except (OSError, subprocess.CalledProcessError):
            if force or not have:
                raise                                      # nothing usable: report the build failure
    gen = os.path.join(ROOT, "tools", "libbscgen.so")
    if force or not os.path.exists(gen):
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-fPIC", "-shared", "-o", gen, os.path.join(ROOT, "tools", "bscgen.c")])


class Gen:
    """Deterministic synthetic inputs."""

    def __init__(self):
        build()
        g = ctypes.CDLL(os.path.join(ROOT, "tools", "libbscgen.so"))
        for f in (g.bscgen_rand, g.bscgen_skew, g.bscgen_text):
            f.argtypes = [ctypes.c_uint64, vp, ctypes.c_size_t]
            f.restype = None
        g.bscgen_adler32.argtypes = [vp, ctypes.c_size_t]
        g.bscgen_adler32.restype = ctypes.c_uint32
        self._g = g

    def _run(self, f, seed, n):
        a = np.empty(n, dtype=np.uint8)
        f(seed, a.ctypes.data, n)
        return a

    def rand(self, seed, n):
        return self._run(self._g.bscgen_rand, seed, n)

    def skew(self, seed, n):
        return self._run(self._g.bscgen_skew, seed, n)

    def text(self, seed, n):
        return self._run(self._g.bscgen_text, seed, n)

    def adler32(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        return int(self._g.bscgen_adler32(a.ctypes.data, a.size))


class _Api:
    """Common numpy-level API over a C library exporting libbsc-shaped functions with `prefix`."""

    def _sig(self, name, args, res=ci):
        f = getattr(self.lib, self.prefix + name)
        f.argtypes = args
        f.restype = res
        return f

    def bwt_encode(self, data, aux=True):
        T = np.array(data, dtype=np.uint8, copy=True)
        idx = (ci * 256)()
        ni = ctypes.c_ubyte(0)
        if aux:
            r = self._bwt_encode(T, ctypes.byref(ni), idx)
        else:
            r = self._bwt_encode(T, None, None)
        return r, T, [idx[t] for t in range(ni.value)]

    def bwt_decode(self, L, index, indexes=()):
        T = np.array(L, dtype=np.uint8, copy=True)
        r = self._bwt_decode(T, index, indexes)
        return r, T

    def st_encode(self, data, k):
        T = np.empty(len(data) + 64, dtype=np.uint8)     # reference writes 28 bytes past n
        T[:len(data)] = data
        r = self._st_encode(T, len(data), k)
        return r, T[:len(data)].copy()

    def coder_compress(self, L, coder=1, features=3):
        L = np.ascontiguousarray(L, dtype=np.uint8)
        out = np.empty(L.size + 4096, dtype=np.uint8)
        r = self._coder_compress(_ptr(L), _ptr(out), L.size, coder, features)
        return r, (out[:r].copy() if r > 0 else None)

    def coder_decompress(self, stream, n, coder=1, features=3):
        s = np.empty(len(stream) + 64, dtype=np.uint8)   # decoder may read a few bytes ahead
        s[:len(stream)] = stream
        s[len(stream):] = 0
        out = np.empty(n + 64, dtype=np.uint8)
        r = self._coder_decompress(s, out, coder, features)
        return r, out[:max(r, 0)].copy()

    def compress(self, data, sorter=1, coder=1, features=3):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.empty(data.size + 28 + 64, dtype=np.uint8)
        r = self._compress(data, out, sorter, coder, features)
        return r, (out[:r].copy() if r > 0 else None)

    def decompress(self, block, features=3):
        block = np.ascontiguousarray(block, dtype=np.uint8)
        bs, ds = ci(0), ci(0)
        r = self._block_info(block, ctypes.byref(bs), ctypes.byref(ds))
        if r != 0:
            return r, None
        src = np.zeros(block.size + 64, dtype=np.uint8)
        src[:block.size] = block
        out = np.empty(ds.value + 64, dtype=np.uint8)
        r = self._decompress(src, block.size, out, ds.value, features)
        return r, out[:ds.value].copy()


class Port(_Api):
    """Our C restatement (liboracle.so)."""
    kind = "port"

    def __init__(self):
        build()
        self.lib = ctypes.CDLL(os.path.join(HERE, "liboracle.so"))
        self.prefix = "orc_"
        s = self._sig
        self.f_adler = s("adler32", [vp, ci], ctypes.c_uint32)
        self.f_bwt_enc = s("bwt_encode", [vp, ci, vp, vp])
        self.f_bwt_dec = s("bwt_decode", [vp, ci, ci])
        self.f_st_enc = s("st_encode", [vp, ci, ci])
        self.f_st_dec = s("st_decode", [vp, ci, ci, ci])
        self.f_transform = s("qlfc_transform", [vp, ci, vp, vp])
        self.f_enc_block = s("qlfc_static_encode_block", [vp, vp, ci, ci])
        self.f_dec_block = s("qlfc_static_decode_block", [vp, vp])
        self.f_enc_fast = s("qlfc_fast_encode_block", [vp, vp, ci, ci])
        self.f_dec_fast = s("qlfc_fast_decode_block", [vp, vp])
        self.f_enc_adapt = s("qlfc_adaptive_encode_block", [vp, vp, ci, ci])
        self.f_dec_adapt = s("qlfc_adaptive_decode_block", [vp, vp])
        self.f_lzp_dec = s("lzp_decompress", [vp, vp, ci, ci, ci, ci])
        self.f_split = s("coder_split_blocks", [vp, ci, ci, vp, vp], None)
        self.f_cc = s("coder_compress", [vp, vp, ci, ci, ci])
        self.f_cd = s("coder_decompress", [vp, vp, ci])
        self.f_store = s("store", [vp, vp, ci])
        self.f_comp = s("compress", [vp, vp, ci, ci, ci, ci])
        self.f_info = s("block_info", [vp, ci, vp, vp])
        self.f_decomp = s("decompress", [vp, ci, vp, ci])

    def adler32(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        return int(self.f_adler(_ptr(a), a.size))

    def _bwt_encode(self, T, ni, idx):
        return self.f_bwt_enc(_ptr(T), T.size, ctypes.cast(ni, vp) if ni is not None else None, ctypes.cast(idx, vp) if idx is not None else None)

    def _bwt_decode(self, T, index, indexes):
        return self.f_bwt_dec(_ptr(T), T.size, index)

    def _st_encode(self, T, n, k):
        return self.f_st_enc(_ptr(T), n, k)

    def st_decode(self, L, k, index):
        T = np.array(L, dtype=np.uint8, copy=True)
        r = self.f_st_dec(_ptr(T), T.size, k, index)
        return r, T

    def transform(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        ranks = np.empty(data.size, dtype=np.uint8)
        mtf = np.empty(256, dtype=np.uint8)
        R = self.f_transform(_ptr(data), data.size, _ptr(ranks), _ptr(mtf))
        return ranks[:R].copy(), mtf

    def encode_block(self, data, out_size=None, coder=1):
        """One QLFC stream (bsc_qlfc_{static,adaptive,fast}_encode_block); coder 1 static, 2 adaptive, 3 fast."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out_size = data.size if out_size is None else out_size
        out = np.empty(data.size + 4096, dtype=np.uint8)
        f = {1: self.f_enc_block, 2: self.f_enc_adapt, 3: self.f_enc_fast}[coder]
        r = f(_ptr(data), _ptr(out), data.size, out_size)
        return r, (out[:r].copy() if r > 0 else None)

    def decode_block(self, stream, n, coder=1):
        s = np.zeros(len(stream) + 64, dtype=np.uint8)
        s[:len(stream)] = stream
        out = np.empty(n + 64, dtype=np.uint8)
        f = {1: self.f_dec_block, 2: self.f_dec_adapt, 3: self.f_dec_fast}[coder]
        r = f(_ptr(s), _ptr(out))
        return r, out[:max(r, 0)].copy()

    def lzp_decompress(self, stream, out_cap, lzp_hash=15, lzp_min=128):
        s = np.ascontiguousarray(stream, dtype=np.uint8)
        out = np.full(out_cap + 64, 0xAA, dtype=np.uint8)
        r = self.f_lzp_dec(_ptr(s), _ptr(out), s.size, out_cap, lzp_hash, lzp_min)
        assert np.all(out[out_cap:] == 0xAA)
        return r, out[:max(r, 0)].copy()

    def split_blocks(self, data, nblocks):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        st, sz = (ci * 8)(), (ci * 8)()
        self.f_split(_ptr(data), data.size, nblocks, ctypes.cast(st, vp), ctypes.cast(sz, vp))
        return list(st[:nblocks]), list(sz[:nblocks])

    def _coder_compress(self, i, o, n, coder, features):
        return self.f_cc(i, o, n, coder, features)

    def _coder_decompress(self, s, out, coder, features):
        return self.f_cd(_ptr(s), _ptr(out), coder)

    def store(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.empty(data.size + 28, dtype=np.uint8)
        r = self.f_store(_ptr(data), _ptr(out), data.size)
        return r, out

    def _compress(self, data, out, sorter, coder, features):
        return self.f_comp(_ptr(data), _ptr(out), data.size, sorter, coder, features)

    def _block_info(self, block, bs, ds):
        return self.f_info(_ptr(block), block.size, ctypes.cast(bs, vp), ctypes.cast(ds, vp))

    def _decompress(self, src, n, out, outn, features):
        return self.f_decomp(_ptr(src), n, _ptr(out), outn)


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libbsc_ref.so"))


class Ref(_Api):
    """The unmodified reference (libbsc 3.3.5, CPU/OpenMP build) -- oracle/_ref/libbsc_ref.so."""
    kind = "reference"

    def __init__(self, init_features=3):
        self.lib = ctypes.CDLL(os.path.join(HERE, "_ref", "libbsc_ref.so"))
        self.prefix = "bsc_"
        s = self._sig
        s("init", [ci])(init_features)
        self.f_adler = s("adler32", [vp, ci, ci], ctypes.c_uint32)
        self.f_bwt_enc = s("bwt_encode", [vp, ci, vp, vp, ci])
        self.f_bwt_dec = s("bwt_decode", [vp, ci, ci, ctypes.c_ubyte, vp, ci])
        self.f_st_enc = s("st_encode", [vp, ci, ci, ci])
        self.f_st_dec = s("st_decode", [vp, ci, ci, ci, ci])
        self.f_enc_block = s("qlfc_static_encode_block", [vp, vp, ci, ci])
        self.f_dec_block = s("qlfc_static_decode_block", [vp, vp])
        self.f_enc_fast = s("qlfc_fast_encode_block", [vp, vp, ci, ci])
        self.f_dec_fast = s("qlfc_fast_decode_block", [vp, vp])
        self.f_enc_adapt = s("qlfc_adaptive_encode_block", [vp, vp, ci, ci])
        self.f_dec_adapt = s("qlfc_adaptive_decode_block", [vp, vp])
        self.f_cc = s("coder_compress", [vp, vp, ci, ci, ci])
        self.f_cd = s("coder_decompress", [vp, vp, ci, ci])
        self.f_store = s("store", [vp, vp, ci, ci])
        self.f_comp = s("compress", [vp, vp, ci, ci, ci, ci, ci, ci])
        self.f_info = s("block_info", [vp, ci, vp, vp, ci])
        self.f_decomp = s("decompress", [vp, ci, vp, ci, ci])
        self.features = 3

    def adler32(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        return int(self.f_adler(_ptr(a), a.size, 0))

    def _bwt_encode(self, T, ni, idx):
        return self.f_bwt_enc(_ptr(T), T.size, ctypes.cast(ni, vp) if ni is not None else None, ctypes.cast(idx, vp) if idx is not None else None, self.features)

    def _bwt_decode(self, T, index, indexes):
        arr = (ci * 256)(*indexes)
        return self.f_bwt_dec(_ptr(T), T.size, index, len(indexes), ctypes.cast(arr, vp), self.features)

    def _st_encode(self, T, n, k):
        return self.f_st_enc(_ptr(T), n, k, self.features)

    def st_decode(self, L, k, index):
        T = np.array(L, dtype=np.uint8, copy=True)
        r = self.f_st_dec(_ptr(T), T.size, k, index, self.features)
        return r, T

    def encode_block(self, data, out_size=None, coder=1):
        """One QLFC stream (bsc_qlfc_{static,adaptive,fast}_encode_block); coder 1 static, 2 adaptive, 3 fast."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out_size = data.size if out_size is None else out_size
        out = np.empty(data.size + 4096, dtype=np.uint8)
        f = {1: self.f_enc_block, 2: self.f_enc_adapt, 3: self.f_enc_fast}[coder]
        r = f(_ptr(data), _ptr(out), data.size, out_size)
        return r, (out[:r].copy() if r > 0 else None)

    def decode_block(self, stream, n, coder=1):
        s = np.zeros(len(stream) + 64, dtype=np.uint8)
        s[:len(stream)] = stream
        out = np.empty(n + 64, dtype=np.uint8)
        f = {1: self.f_dec_block, 2: self.f_dec_adapt, 3: self.f_dec_fast}[coder]
        r = f(_ptr(s), _ptr(out))
        return r, out[:max(r, 0)].copy()

    def _coder_compress(self, i, o, n, coder, features):
        return self.f_cc(i, o, n, coder, features)

    def _coder_decompress(self, s, out, coder, features):
        return self.f_cd(_ptr(s), _ptr(out), coder, features)

    def store(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.empty(data.size + 28, dtype=np.uint8)
        r = self.f_store(_ptr(data), _ptr(out), data.size, 0)
        return r, out

    def _compress(self, data, out, sorter, coder, features):
        return self.f_comp(_ptr(data), _ptr(out), data.size, 0, 0, sorter, coder, features)

    def compress_lzp(self, data, lzp_hash=15, lzp_min=128, sorter=1, coder=1, features=3):
        """bsc_compress with the reference's LZP stage on (its default options)."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.empty(data.size + 28 + 64, dtype=np.uint8)
        r = self.f_comp(_ptr(data), _ptr(out), data.size, lzp_hash, lzp_min, sorter, coder, features)
        return r, (out[:r].copy() if r > 0 else None)

    def lzp_compress(self, data, lzp_hash=15, lzp_min=128, features=3):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.empty(data.size + 64, dtype=np.uint8)
        f = self._sig("lzp_compress", [vp, vp, ci, ci, ci, ci])
        r = f(_ptr(data), _ptr(out), data.size, lzp_hash, lzp_min, features)
        return r, (out[:r].copy() if r > 0 else None)

    def _block_info(self, block, bs, ds):
        return self.f_info(_ptr(block), block.size, ctypes.cast(bs, vp), ctypes.cast(ds, vp), 0)

    def _decompress(self, src, n, out, outn, features):
        return self.f_decomp(_ptr(src), n, _ptr(out), outn, features)


def best():
    """The strongest checker available on this machine: the real reference if built, else the port."""
    return Ref() if ref_available() else Port()
