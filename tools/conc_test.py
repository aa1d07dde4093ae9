#!/usr/bin/env python3
"""tools/conc_test.py -- do K blocks in K contexts really overlap on one GPU?
Prints the wall time of K concurrent stage calls for K = 1, 2, 4, 8, 16 (same block each)."""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libbsc_b200
from oracle import pyoracle

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 32
gen = pyoracle.Gen()
L = libbsc_b200.lib()
assert L.bsc_init(3) == 0
n = mib << 20
dev = torch.device("cuda", 0)
text = torch.from_numpy(gen.text(2, n)).to(dev)
KMAX = 16
ctxs = [libbsc_b200.DeviceCtx(0) for _ in range(KMAX)]
ws = int(L.bscb200_workspace_bytes(n, 1))
for c in ctxs:
    assert c.reserve(ws) == 0
work = [text.clone() for _ in range(KMAX)]
outs = [torch.empty(n + 4096, dtype=torch.uint8, device=dev) for _ in range(KMAX)]
back = [torch.empty(n + 64, dtype=torch.uint8, device=dev) for _ in range(KMAX)]
idx = [0] * KMAX
csz = [0] * KMAX
pool = ThreadPoolExecutor(max_workers=KMAX)


def phase(name, fn, K):
    torch.cuda.synchronize()
    t = time.perf_counter()
    list(pool.map(fn, range(K)))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("%-18s K=%2d  %8.1f ms  (%7.1f MB/s aggregate)" % (name, K, dt * 1e3, K * n / 1e6 / dt), flush=True)


def f_bwt(i):
    work[i].copy_(text); torch.cuda.synchronize()
    idx[i] = ctxs[i].bwt_encode(work[i].data_ptr(), n, aux=False)[0]


def f_enc(i):
    csz[i] = ctxs[i].coder_compress(work[i].data_ptr(), outs[i].data_ptr(), n, 1, 3)
    assert csz[i] > 0, csz[i]


def f_dec(i):
    r = ctxs[i].coder_decompress(outs[i].data_ptr(), csz[i], back[i].data_ptr(), n, 1, 3)
    assert r == n, r


def f_unbwt(i):
    r = ctxs[i].bwt_decode(back[i].data_ptr(), n, idx[i])
    assert r == 0


for K in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "1,1,2,4,8,16".split(","))]:
    phase("bwt_encode", f_bwt, K)
    phase("coder_compress", f_enc, K)
    phase("coder_decompress", f_dec, K)
    phase("bwt_decode", f_unbwt, K)
    for i in range(K):
        assert torch.equal(back[i][:n], text), "round trip mismatch"
ctxs[0].set_profile(True)
f_bwt(0); f_enc(0); f_dec(0); f_unbwt(0)
for k, v in sorted(ctxs[0].profile_report().items(), key=lambda kv: -kv[1][1]):
    print("%-20s n=%4d %10.3f ms" % (k, v[0], v[1]))
