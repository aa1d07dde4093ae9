"""tools/sass_t1w.py (single-warp timing model over SASS control bits) keeps parsing what nvcc / cuobjdump produce."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "libbsc_b200", "build", "qlfc.o")


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(OBJ), reason="needs cuobjdump and a built libbsc_b200/build/qlfc.o")
def test_model_reads_the_default_decoder():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_t1w as T
    name, ins = T.disasm(OBJ, "q_decode3ILi1ELb0")                      # the default decoder
    assert "q_decode3" in name and len(ins) > 1500
    assert all(0 <= i.stall <= 15 and 0 <= i.wait < 64 for i in ins)
    bbs = T.blocks(ins)
    assert sum(len(b) for b in bbs) == len(ins)
    loops = [b for b in bbs if b[-1].target is not None and b[-1].target == b[0].addr and len(b) >= 30]
    assert loops, "the decision loops of the decoder end in a branch to their own head"
    t, _, exposed = T.simulate(loops[0])
    assert 40 <= t <= 400 and exposed >= 0                               # a 30-60 instruction decision: tens to a few hundred cycles
