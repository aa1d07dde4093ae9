"""Host-side logic of bench.py that needs no GPU: the block pipeline (task order, workers independent of the blocks per step, per-block
exclusivity for the extra configurations, error propagation) and the reference arm's JSON line on a tiny workload."""
import json
import os
import subprocess
import sys
import threading
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def test_pipeline_runs_every_task_once_per_step():
    seen, lock = [], threading.Lock()

    def task(w, i):
        with lock:
            seen.append((w, i))
    bench.Pipeline(5, 12, task).run(3)                       # more workers than blocks per step
    assert sorted(i for _, i in seen) == sorted(list(range(5)) * 3)
    assert all(0 <= w < 12 for w, _ in seen)
    seen.clear()
    bench.Pipeline(7, 2, task).run(1)
    assert sorted(i for _, i in seen) == list(range(7)) and {w for w, _ in seen} <= {0, 1}


def test_pipeline_block_buffers_are_exclusive():
    busy, clash, lock = set(), [], threading.Lock()

    def task(w, i):
        with lock:
            if i in busy:
                clash.append(i)
            busy.add(i)
        time.sleep(0.002)
        with lock:
            busy.discard(i)
    bench.Pipeline(3, 9, task, block_buffers=True).run(6)
    assert not clash


def test_pipeline_surfaces_the_first_failure():
    def task(w, i):
        if i == 3:
            raise ValueError("block 3")
    with pytest.raises(ValueError):
        bench.Pipeline(8, 4, task).run(2)


def test_reference_arm_line_on_a_tiny_workload():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "librefdrv.so")):
        pytest.skip("oracle/_ref not built on this machine")
    env = dict(os.environ, OMP_NUM_THREADS="1")              # what torchrun exports; the arm must not inherit it
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--blocks", "2", "--block-mib", "1"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-500:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "MB/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == min(2, len(os.sched_getaffinity(0)))
    assert line["e2e"] == {"value": line["value"], "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    st = line["cpu_baseline"]["stages"]                      # per-stage seconds of one block, one thread
    assert all(st[k] > 0 for k in ("bwt_encode_s", "coder_compress_s", "coder_decompress_s", "bwt_decode_s")) and st["compressed_bytes"] > 0
    assert line["config"]["blocks_per_gpu"] == 2 and "arm" in line
