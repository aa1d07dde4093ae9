"""The algebra behind the RANGE3 variant of the static encoder's range warp (libbsc_b200/csrc/qlfc_encoder.cuh,
BSCB200_QENC=2): one multiply-add per record must equal the reference recurrence (rangecoder.h:145-177) in 32-bit
arithmetic, including the addend of `low` and the renormalisation flag."""
import numpy as np


def test_one_multiply_add_equals_reference_recurrence():
    rng = np.random.default_rng(7)
    n = 2_000_000
    # ranges as the coder sees them: anything from 1 .. 2^32-1, with many values just around the 2^16 renormalisation edge
    rangev = np.concatenate([rng.integers(1, 1 << 32, n // 2, dtype=np.uint64), rng.integers(1, 1 << 17, n // 2, dtype=np.uint64)]).astype(np.uint32)
    p = rng.integers(1, 4096, n, dtype=np.uint64).astype(np.uint32)
    bit = rng.integers(0, 2, n).astype(bool)

    sh = rangev < 0x10000
    # reference form
    r0 = np.where(sh, rangev << np.uint32(16), rangev)
    r = (r0 >> np.uint32(12)) * p                                  # uint32 wrap like the device
    ref_range = np.where(bit, r0 - r, r)
    ref_add = np.where(bit, r, np.uint32(0))
    # RANGE3 form
    mul = np.where(bit, np.uint32(0) - p, p)
    keep = np.where(bit, np.uint32(0xFFFFFFFF), np.uint32(0))
    rin = np.where(sh, rangev << np.uint32(4), rangev >> np.uint32(12))
    rs = np.where(sh, rangev << np.uint32(16), rangev)
    new_add = (rin * p) & keep
    new_range = rin * mul + (rs & keep)
    assert np.array_equal(new_range, ref_range)
    assert np.array_equal(new_add, ref_add)
