#!/usr/bin/env python3
"""tools/time_ref_cuda.py -- the reference's OWN CUDA path (libcubwt forward / inverse BWT, st.cu forward ST) against ours, same box,
same host-pointer entry points (bsc_bwt_encode / bsc_bwt_decode / bsc_st_encode: H2D + kernels + D2H inside both timings).

The reference side is oracle/_ref/libbsc_refcuda.so: the unmodified reference compiled by oracle/Makefile with -DLIBBSC_CUDA_SUPPORT for
sm_100 (SURVEY 8c "second baseline"; rows a3 / a6 / a8).  Both sides get `reps` warm calls; best and median wall times are reported,
plus our own kernel-only time (CUDA events around every launch of the stage).  Outputs are compared byte for byte.
    python tools/time_ref_cuda.py [reps] > gpurun_out/ref_cuda.json"""
import ctypes
import json
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libbsc_b200
from oracle import pyoracle

vp, ci = ctypes.c_void_p, ctypes.c_int
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
gen = pyoracle.Gen()
R = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libbsc_refcuda.so"))
FEAT = 1 | 2 | 8                                         # FASTMODE | MULTITHREADING | CUDA (bsc.cpp -G)
assert R.bsc_init(FEAT) == 0
R.bsc_bwt_encode.argtypes = [vp, ci, vp, vp, ci]
R.bsc_bwt_decode.argtypes = [vp, ci, ci, ctypes.c_ubyte, vp, ci]
R.bsc_st_encode.argtypes = [vp, ci, ci, ci]
B = libbsc_b200.Bsc(features=3)
L = B.lib


def timed(fn, reps):
    ts = []
    out = None
    for _ in range(reps + 1):                            # first call warms allocations / module load
        t0 = time.perf_counter(); out = fn(); ts.append(time.perf_counter() - t0)
    ts = ts[1:]
    return out, min(ts) * 1e3, float(np.median(ts)) * 1e3


def ours_kernel_ms(stage, data, **kw):
    """kernel-only time of our stage: device-resident call on a profiled context"""
    import torch
    dev = torch.device("cuda", 0)
    n = data.size
    d = torch.from_numpy(data).to(dev)
    pad = torch.empty(n + 64, dtype=torch.uint8, device=dev); pad[:n] = d
    ctx = libbsc_b200.DeviceCtx(0)
    assert ctx.reserve(int(L.bscb200_workspace_bytes(n, kw.get("sorter", 1)))) == 0
    best = None
    for rep in range(3):
        pad[:n] = d
        ctx.set_profile(True)
        if stage == "bwt_encode":
            ctx.bwt_encode(pad.data_ptr(), n)
        elif stage == "bwt_decode":
            ctx.bwt_decode(pad.data_ptr(), n, kw["index"])
        elif stage == "st_encode":
            ctx.st_encode(pad.data_ptr(), n, kw["k"])
        elif stage == "st_decode":
            ctx.st_decode(pad.data_ptr(), n, kw["k"], kw["index"])
        rep_ms = sum(ms for (_, ms, _) in ctx.profile_report().values())
        ctx.set_profile(False)
        best = rep_ms if best is None else min(best, rep_ms)
    ctx.close()
    return best


rows = []
text = gen.text(2, 64 << 20)
skew = gen.skew(3, 32 << 20)

# ---- forward BWT, 64 MiB text (a3: libcubwt_bwt_aux through bsc_bwt_encode) ----
def ref_bwt_enc():
    T = text.copy(); ni = ctypes.c_ubyte(0); idx = (ci * 256)()
    r = R.bsc_bwt_encode(T.ctypes.data, T.size, ctypes.byref(ni), idx, FEAT)
    return r, T
(r_ref, L_ref), rb, rm = timed(ref_bwt_enc, reps)
(r_our, L_our, _), ob, om = timed(lambda: B.bwt_encode(text), reps)
assert r_ref == r_our and np.array_equal(L_ref, L_our), "forward BWT differs from the reference's CUDA path"
rows.append({"stage": "bsc_bwt_encode 64 MiB G_text", "ref_cuda_ms_best": rb, "ref_cuda_ms_median": rm, "ours_ms_best": ob, "ours_ms_median": om,
             "ours_kernels_only_ms": ours_kernel_ms("bwt_encode", text), "bit_identical": True})

# ---- inverse BWT (a6: libcubwt_unbwt through bsc_bwt_decode) ----
def ref_bwt_dec():
    T = L_ref.copy()
    r = R.bsc_bwt_decode(T.ctypes.data, T.size, r_ref, 0, None, FEAT)
    return r, T
(q_ref, T_ref), rb, rm = timed(ref_bwt_dec, reps)
(q_our, T_our), ob, om = timed(lambda: B.bwt_decode(L_ref, r_ref), reps)
assert q_ref == 0 and q_our == 0 and np.array_equal(T_ref, text) and np.array_equal(T_our, text)
rows.append({"stage": "bsc_bwt_decode 64 MiB G_text", "ref_cuda_ms_best": rb, "ref_cuda_ms_median": rm, "ours_ms_best": ob, "ours_ms_median": om,
             "ours_kernels_only_ms": ours_kernel_ms("bwt_decode", L_ref, index=r_ref), "bit_identical": True})

# ---- forward ST (a8: bsc_st_encode_cuda, k = 5..8) on 32 MiB skew ----
st_out = {}
for k in (5, 6, 7, 8):
    def ref_st():
        T = np.empty(skew.size + 64, dtype=np.uint8); T[:skew.size] = skew
        r = R.bsc_st_encode(T.ctypes.data, skew.size, k, FEAT)
        return r, T[:skew.size]
    (i_ref, S_ref), rb, rm = timed(ref_st, reps)
    (i_our, S_our), ob, om = timed(lambda: B.st_encode(skew, k), reps)
    assert i_ref == i_our and np.array_equal(S_ref, S_our), "ST%d differs from the reference's CUDA path" % k
    st_out[k] = (i_our, S_our.copy())
    rows.append({"stage": "bsc_st_encode k=%d 32 MiB G_skew" % k, "ref_cuda_ms_best": rb, "ref_cuda_ms_median": rm, "ours_ms_best": ob, "ours_ms_median": om,
                 "ours_kernels_only_ms": ours_kernel_ms("st_encode", skew, k=k, sorter=k), "bit_identical": True})

# ---- inverse ST: the reference has no GPU path (st.cpp:1491, CPU only) ----
if hasattr(B, "st_decode"):
    R.bsc_st_decode.argtypes = [vp, ci, ci, ci, ci]
    i6, S6 = st_out[6]
    def ref_unst():
        T = S6.copy(); r = R.bsc_st_decode(T.ctypes.data, T.size, 6, i6, FEAT); return r, T
    (q_ref, U_ref), rb, rm = timed(ref_unst, 1)
    (q_our, U_our), ob, om = timed(lambda: B.st_decode(S6, 6, i6), reps)
    assert q_ref == 0 and q_our == 0 and np.array_equal(U_ref, skew) and np.array_equal(U_our, skew)
    rows.append({"stage": "bsc_st_decode k=6 32 MiB G_skew (reference: CPU, OpenMP)", "ref_cuda_ms_best": rb, "ref_cuda_ms_median": rm, "ours_ms_best": ob, "ours_ms_median": om,
                 "ours_kernels_only_ms": ours_kernel_ms("st_decode", S6, k=6, index=i6, sorter=6), "bit_identical": True})

for r in rows:
    r["speedup_best"] = round(r["ref_cuda_ms_best"] / r["ours_ms_best"], 2)
    sys.stderr.write("%-58s ref %9.1f ms   ours %9.1f ms (kernels %8.1f)   x%.2f\n" % (r["stage"], r["ref_cuda_ms_best"], r["ours_ms_best"], r["ours_kernels_only_ms"], r["speedup_best"]))
print(json.dumps({"what": "reference CUDA build (libcubwt 1.6.1 / st.cu, sm_100 SASS) vs libbsc_b200, host-pointer calls, wall clock incl. H2D/D2H", "reps": reps, "rows": rows}, indent=1))
