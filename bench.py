#!/usr/bin/env python3
"""bench.py -- compress+decompress throughput of the block-sorting hot path on B200.

One "step" = one pass of the hot path over this rank's batch of synthetic blocks: every block is
compressed (adler32 -> forward BWT -> QLFC static coder -> framing) and then decompressed (QLFC
decode -> inverse BWT -> adler32 check), all on the GPU through the C ABI of libbsc_b200.so.

  value : whole-job MB/s (10^6 uncompressed bytes, CLI convention bsc.cpp:427) with the blocks already
          resident in HBM (bscb200_compress_device / bscb200_decompress_device), CUDA-event timed.
  e2e   : the same batch through the reference-facing host-pointer entry points bsc_compress /
          bsc_decompress, from pinned HOST buffers to pinned HOST buffers (H2D/D2H inside the timed region).
  roofline / kernels : per-kernel CUDA-event durations captured during the timed steps.
  cpu_baseline / --impl reference : the UNMODIFIED reference (oracle/_ref) driven like its CLI
          (oracle/ref_driver.c) on the box's host cores.

Workload (BASELINE.json config C3, per GPU): G_text(seed 2 + rank), 18 blocks of 64 MiB, sorter BWT,
coder QLFC static, LZP off.  18 blocks = 144 coder streams = one per SM on 144 of the 148 SMs (the coder stage is one
SM per stream).  Weak scaling: every rank processes its own 1.125 GiB.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

# One hardware work queue per stream: with the default of 8, streams 9..16 alias onto the queues of
# streams 1..8 and a long-running coder kernel falsely serialises the other block (measured: 2x).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "compress+decompress MB/s on 64 MB blocks"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--blocks", type=int, default=18, help="blocks per GPU (18 x 8 coder streams = 144 of the 148 SMs, one stream per SM)")
    ap.add_argument("--block-mib", type=int, default=64)
    ap.add_argument("--workers", type=int, default=0, help="concurrent blocks per GPU (0 = all)")
    ap.add_argument("--decode-workers", type=int, default=0,
                    help="separate, smaller contexts for decompression (0 = reuse the compression contexts); e.g. --blocks 36 --workers 18 "
                         "--decode-workers 36 keeps 288 decoder streams in flight where 36 full contexts (4.6 GiB each) would not fit")
    ap.add_argument("--sorter", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def make_blocks(gen, seed, nblocks, block_bytes):
    data = gen.text(seed, nblocks * block_bytes)
    return [data[i * block_bytes:(i + 1) * block_bytes] for i in range(nblocks)]


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the unmodified reference on the host cores, CLI-style block loop."""
    from oracle import pyoracle
    if rank != 0:
        return
    # Same workload as the GPU arm (world x blocks-per-GPU blocks of the same size), bounded to ONE block per host thread:
    # the reference's throughput saturates there (one block per OpenMP thread, bsc.cpp:184-199), more blocks only lengthen the step.
    threads = len(os.sched_getaffinity(0))
    sample = max(1, min(world * args.blocks, threads))
    info = reference_timing(args, steps=args.steps, warmup=args.warmup, sample_blocks=sample)
    line = {"metric": METRIC, "value": info["value"], "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "impl": "reference",
            "config": workload_config(args, extra={"host_threads": info["cores"]}),
            "cpu_baseline": {"value": info["value"], "unit": "MB/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": info["value"], "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "compress_MBps": info["compress_MBps"], "decompress_MBps": info["decompress_MBps"]}
    print(json.dumps(line), flush=True)


def reference_timing(args, steps, warmup, sample_blocks=None):
    from oracle import pyoracle
    gen = pyoracle.Gen()
    nb = sample_blocks or args.blocks
    bb = args.block_mib << 20
    blocks = make_blocks(gen, 2, nb, bb)
    drv_path = os.path.join(ROOT, "oracle", "_ref", "librefdrv.so")
    if os.path.exists(drv_path):
        return _reference_native(drv_path, blocks, steps, warmup, args.sorter)
    # no compiled reference on this machine: time the oracle port, one block per Python thread
    port = pyoracle.Port()
    t0 = time.perf_counter()
    outs = [port.compress(b, args.sorter, 1, 3) for b in blocks[:1]]
    t1 = time.perf_counter()
    _ = [port.decompress(o[1]) for o in outs]
    t2 = time.perf_counter()
    mb = bb / 1e6
    return {"value": mb / (t2 - t0), "ms_per_step": (t2 - t0) * 1e3, "cores": 1, "kind": "port", "sample": "1 block of %d MiB, single thread" % args.block_mib,
            "compress_MBps": mb / (t1 - t0), "decompress_MBps": mb / (t2 - t1)}


def _reference_native(drv_path, blocks, steps, warmup, sorter):
    vp, ci = ctypes.c_void_p, ctypes.c_int
    os.environ.pop("OMP_NUM_THREADS", None)                  # torchrun exports OMP_NUM_THREADS=1 to its ranks; the baseline uses the whole host
    d = ctypes.CDLL(drv_path)
    d.refdrv_init()
    d.refdrv_set_threads(len(os.sched_getaffinity(0)))
    d.refdrv_set_nested(1 if os.environ.get("BSCB200_REF_NESTED") == "1" else 0)   # stock CLI behaviour by default (see oracle/ref_driver.c)
    nb = len(blocks)
    PP = ctypes.c_void_p * nb
    II = ci * nb
    outs = [np.empty(b.size + 28 + 64, dtype=np.uint8) for b in blocks]
    backs = [np.empty(b.size + 64, dtype=np.uint8) for b in blocks]
    inp = PP(*[b.ctypes.data for b in blocks]); sizes = II(*[b.size for b in blocks])
    outp = PP(*[o.ctypes.data for o in outs]); outsz = II()
    backp = PP(*[o.ctypes.data for o in backs]); res = II()
    d.refdrv_compress.argtypes = [vp, vp, ci, vp, vp, ci, ci]
    d.refdrv_decompress.argtypes = [vp, vp, ci, vp, vp, vp]
    tc, td, T = [], [], 1
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        T = d.refdrv_compress(inp, sizes, nb, outp, outsz, sorter, 1)
        t1 = time.perf_counter()
        d.refdrv_decompress(outp, outsz, nb, backp, sizes, res)
        t2 = time.perf_counter()
        assert all(r == 0 for r in res), "reference decompress failed"
        if it >= warmup:
            tc.append(t1 - t0); td.append(t2 - t1)
    for b, o in zip(blocks, backs):
        assert np.array_equal(b, o[:b.size])
    mb = sum(b.size for b in blocks) / 1e6
    c, dd = float(np.mean(tc)), float(np.mean(td))
    return {"value": mb / (c + dd), "ms_per_step": (c + dd) * 1e3, "cores": int(T), "kind": "reference",
            "sample": "%d blocks x %d MiB G_text, CLI block loop (bsc.cpp:184-199), %d OpenMP threads of %d host threads" % (nb, blocks[0].size >> 20, T, d.refdrv_max_threads()),
            "compress_MBps": mb / c, "decompress_MBps": mb / dd, "compressed_bytes": int(sum(outsz))}


def workload_config(args, extra=None):
    c = {"workload": "C3: G_text(seed 2+rank) %d x %d MiB blocks per GPU, BWT + QLFC static, LZP off (-b%d -m0 -e1 -p)" % (args.blocks, args.block_mib, args.block_mib),
         "blocks_per_gpu": args.blocks, "block_bytes": args.block_mib << 20, "sorter": args.sorter, "coder": 1,
         "l2": "inputs_larger_than_L2 (%d MiB per step vs 126 MB L2)" % (args.blocks * args.block_mib)}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import torch
    import libbsc_b200
    from oracle import pyoracle            # generators only (tools/libbscgen.so) + the cpu_baseline leg

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = libbsc_b200.lib()
    assert L.bsc_init(3) == 0, "no usable CUDA device (libbsc_b200 has no CPU path)"

    gen = pyoracle.Gen()
    nb, bb = args.blocks, args.block_mib << 20
    host_blocks = make_blocks(gen, 2 + rank, nb, bb)
    workers = args.workers or nb

    # ---- device-resident leg ------------------------------------------------------------------
    d_in = [torch.from_numpy(b).to(dev) for b in host_blocks]
    d_cmp = [torch.empty(bb + 28 + 64, dtype=torch.uint8, device=dev) for _ in range(nb)]
    d_back = [torch.empty(bb + 64, dtype=torch.uint8, device=dev) for _ in range(nb)]
    ctxs = [libbsc_b200.DeviceCtx(local_rank) for _ in range(workers)]
    ws = int(L.bscb200_workspace_bytes(bb, args.sorter))
    for c in ctxs:
        assert c.reserve(ws) == 0, "workspace allocation failed"
    dworkers = args.decode_workers or workers
    dctxs = ctxs
    if args.decode_workers:
        dctxs = [libbsc_b200.DeviceCtx(local_rank) for _ in range(dworkers)]
        for c in dctxs:
            assert c.reserve(int(L.bscb200_workspace_bytes_decode(bb))) == 0, "decode workspace allocation failed"
    allctx = ctxs + ([] if dctxs is ctxs else dctxs)
    ctx_lock = {id(c): threading.Lock() for c in allctx}       # one block at a time per context (blocks > contexts share them)
    csize = [0] * nb
    pool = ThreadPoolExecutor(max_workers=max(workers, dworkers))

    def dev_compress(i):
        torch.cuda.set_device(local_rank)
        c = ctxs[i % workers]
        # +4: payload (offset 28) 16-byte aligned for the vectorised device adler32
        with ctx_lock[id(c)]:
            r = c.compress(d_in[i].data_ptr(), d_cmp[i].data_ptr() + 4, bb, args.sorter, 1, 3)
        assert r > 0, "compress failed: %d" % r
        csize[i] = r

    def dev_decompress(i):
        torch.cuda.set_device(local_rank)
        c = dctxs[i % dworkers]
        with ctx_lock[id(c)]:
            r = c.decompress(d_cmp[i].data_ptr() + 4, csize[i], d_back[i].data_ptr(), bb, 3)
        assert r == 0, "decompress failed: %d" % r

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_phase(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        list(pool.map(fn, range(nb)))
        torch.cuda.synchronize()
        e1.record(); e1.synchronize()
        return e0.elapsed_time(e1)

    for _ in range(args.warmup):
        timed_phase(dev_compress); timed_phase(dev_decompress)
    for c in allctx:
        c.set_profile(True)
    launches0 = sum(c.launches() for c in allctx)
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    t_c = t_d = 0.0
    for _ in range(args.steps):
        t_c += timed_phase(dev_compress)
        t_d += timed_phase(dev_decompress)
    barrier()
    clocks = sampler.stop()
    launches = sum(c.launches() for c in allctx) - launches0
    for i in range(nb):
        assert torch.equal(d_back[i][:bb], d_in[i]), "round trip mismatch in block %d" % i
    # per-kernel CUDA-event timings gathered during the timed steps
    kern = {}
    for c in allctx:
        for name, (cnt, ms, by) in c.profile_report().items():
            a = kern.setdefault(name, [0, 0.0, 0.0]); a[0] += cnt; a[1] += ms; a[2] += by
        c.set_profile(False)

    # standalone pass: ONE block at a time on one stream, so that the per-launch durations of the
    # bandwidth-bound kernels are not stretched by 15 other blocks sharing HBM (roofline leg)
    alone = {}
    ctxs[0].set_profile(True)
    for i in range(min(2, nb)):
        csz_i = ctxs[0].compress(d_in[i].data_ptr(), d_cmp[i].data_ptr() + 4, bb, args.sorter, 1, 3)
        assert csz_i == csize[i]
        assert ctxs[0].decompress(d_cmp[i].data_ptr() + 4, csz_i, d_back[i].data_ptr(), bb, 3) == 0
    for name, (cnt, ms, by) in ctxs[0].profile_report().items():
        alone[name] = [cnt, ms, by]
    ctxs[0].set_profile(False)

    ms_c, ms_d = t_c / args.steps, t_d / args.steps
    if dist is not None:
        t = torch.tensor([ms_c, ms_d], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_c, ms_d = float(t[0]), float(t[1])
    total_mb = world * nb * bb / 1e6
    value = total_mb / ((ms_c + ms_d) / 1e3)

    # ---- end-to-end leg: pinned host buffers through bsc_compress / bsc_decompress ---------------
    e2e = None
    comp_bytes = int(sum(csize))
    if not args.no_e2e:
        for c in allctx:
            c.close()
        del d_cmp, d_back
        torch.cuda.empty_cache()
        h_in = [torch.from_numpy(b).pin_memory() for b in host_blocks]
        h_cmp = [torch.empty(bb + 28 + 64, dtype=torch.uint8).pin_memory() for _ in range(nb)]
        h_back = [torch.empty(bb + 64, dtype=torch.uint8).pin_memory() for _ in range(nb)]
        hsize = [0] * nb

        def host_compress(i):
            torch.cuda.set_device(local_rank)
            r = L.bsc_compress(h_in[i].data_ptr(), h_cmp[i].data_ptr(), bb, 0, 0, args.sorter, 1, 3)
            assert r > 0, "bsc_compress failed: %d" % r
            hsize[i] = r

        def host_decompress(i):
            torch.cuda.set_device(local_rank)
            r = L.bsc_decompress(h_cmp[i].data_ptr(), hsize[i], h_back[i].data_ptr(), bb, 3)
            assert r == 0, "bsc_decompress failed: %d" % r

        for _ in range(max(1, args.warmup - 1)):
            timed_phase(host_compress); timed_phase(host_decompress)
        l0 = int(L.bscb200_total_kernel_launches())
        barrier()
        e_c = e_d = 0.0
        for _ in range(args.steps):
            e_c += timed_phase(host_compress)
            e_d += timed_phase(host_decompress)
        barrier()
        launches += int(L.bscb200_total_kernel_launches()) - l0
        for i in range(nb):
            assert torch.equal(h_back[i][:bb], h_in[i]), "e2e round trip mismatch in block %d" % i
        e_c, e_d = e_c / args.steps, e_d / args.steps
        if dist is not None:
            t = torch.tensor([e_c, e_d], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_c, e_d = float(t[0]), float(t[1])
        e2e = {"value": total_mb / ((e_c + e_d) / 1e3), "unit": "MB/s",
               "h2d_bytes_per_step": nb * bb + int(sum(hsize)), "d2h_bytes_per_step": int(sum(hsize)) + nb * bb,
               "compress_MBps": total_mb / (e_c / 1e3), "decompress_MBps": total_mb / (e_d / 1e3), "ms_per_step": e_c + e_d}

    if rank != 0:
        return

    # ---- roofline of the dominant kernel + per-kernel table ------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    tot_ms = sum(v[1] for v in kern.values()) or 1.0
    table = []
    for name, (cnt, ms, by) in sorted(kern.items(), key=lambda kv: -kv[1][1]):
        row = {"kernel": name, "launches": cnt, "ms_total": round(ms, 3), "share": round(ms / tot_ms, 4)}
        if by > 0:
            row["algorithmic_GBps"] = round(by / 1e9 / (ms / 1e3), 2); row["frac_of_peak"] = round(by / 1e9 / (ms / 1e3) / peak, 5)
        table.append(row)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except Exception:
        pass
    dom = table[0] if table else {"kernel": None}
    dom_bytes = kern[dom["kernel"]][2] if table else 0
    dom_ms = kern[dom["kernel"]][1] if table else 1
    dom_cnt = kern[dom["kernel"]][0] if table else 1
    achieved = dom_bytes / 1e9 / (dom_ms / 1e3) if dom_bytes else 0.0
    roofline = {"kernel": dom["kernel"], "bound": "hbm", "achieved": round(achieved, 3), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 6),
                "traffic": (traffic or {}).get(dom["kernel"]), "peak_source": peak_src,
                "bytes_per_launch": dom_bytes / max(dom_cnt, 1), "avg_launch_ms": dom_ms / max(dom_cnt, 1), "share_of_kernel_time": dom.get("share"),
                "note": "concurrent streams: kernel durations overlap, shares are of summed kernel time"}

    # HBM-bound kernels, timed alone: achieved algorithmic GB/s against the measured copy peak
    alone_table = []
    for name, (cnt, ms, by) in sorted(alone.items(), key=lambda kv: -kv[1][1]):
        row = {"kernel": name, "launches": cnt, "ms_total": round(ms, 3)}
        if by > 0 and ms > 0:
            row["algorithmic_GBps"] = round(by / 1e9 / (ms / 1e3), 1); row["frac_of_peak"] = round(by / 1e9 / (ms / 1e3) / peak, 4)
        alone_table.append(row)
    hbm_rows = [r for r in alone_table if r["kernel"] in ("rs_onesweep", "unbwt_lf", "rs_hist1", "adler_partial") and "algorithmic_GBps" in r]
    roofline_hbm = None
    if hbm_rows:
        r0 = hbm_rows[0]; cnt0, ms0, by0 = alone[r0["kernel"]]
        roofline_hbm = {"kernel": r0["kernel"], "bound": "hbm", "achieved": r0["algorithmic_GBps"], "peak": peak, "unit": "GB/s", "frac": r0["frac_of_peak"],
                        "traffic": (traffic or {}).get(r0["kernel"]), "bytes_per_launch": by0 / max(cnt0, 1), "avg_launch_ms": ms0 / max(cnt0, 1),
                        "how": "CUDA events around every launch, one block at a time on one stream (2 blocks), after the timed steps"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            info = reference_timing(args, steps=1, warmup=0)
            cpu = {"value": info["value"], "unit": "MB/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"],
                   "compress_MBps": info["compress_MBps"], "decompress_MBps": info["decompress_MBps"]}
        except Exception as ex:      # never lose the GPU line because the baseline leg failed
            cpu = {"value": None, "unit": "MB/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (ex,)}

    line = {"metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_c + ms_d, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, {"concurrent_blocks_per_gpu": workers, "concurrent_decode_blocks_per_gpu": dworkers, "parallelism": "blocks round-robin over %d GPU(s), no collective" % world}),
            "compress_MBps": total_mb / (ms_c / 1e3), "decompress_MBps": total_mb / (ms_d / 1e3),
            "compressed_bytes_rank0": comp_bytes, "ratio": comp_bytes / float(nb * bb),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_hbm_kernel": roofline_hbm,
            "kernels": table, "kernels_standalone": alone_table, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
