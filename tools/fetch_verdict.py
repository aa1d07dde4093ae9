#!/usr/bin/env python3
"""tools/fetch_verdict.py [warp_latency output file]  (default gpurun_out/r2_first_call.log)

Reads the measured cycles of tools/warp_latency.cu and sets them against the single-warp model
(profiles/r1i_model_predictions.txt): per case measured / modelled, and for the control-flow cases the extra cycles per taken
branch as a function of the loop's code footprint -- the instruction-fetch question of DESIGN.md 4.5."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LABEL = {"lone warp, 1 IMAD chain": 100, "lone warp, 2 independent": 102, "lone warp, 4 independent": 104, "lone warp, 8 independent": 108,
         "lone warp, 64 independent IMAD": 110, "lone warp, 32 IMAD + 32 LOP3": 111, "lone warp, 2048 independent": 112,
         "8 uniform forward branches, all TAKEN": 120, "8 uniform forward branches, none taken": 121,
         "32 taken forward branches, footprint 9.5 KB": 124, "32 taken forward branches, footprint 34 KB": 125, "64 taken forward branches, footprint 68 KB": 126,
         "IMAD -> SHF -> ISETP -> taken BRA": 122, "IMAD -> SHF -> ISETP -> predicated": 123, "LDS.U16 pointer chase": 130,
         "IMAD + SHF (counter move)": 5, "LDS (pointer chase)": 3, "__shfl_sync": 0, "__ballot_sync": 4, "__reduce_max_sync": 2}
BRANCHES = {120: (8, "1.5 KB"), 124: (32, "9.5 KB"), 125: (32, "34 KB"), 126: (64, "68 KB")}


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2_first_call.log")
    model = {}
    for line in open(os.path.join(ROOT, "profiles", "r1i_model_predictions.txt")):
        m = re.match(r"case\s+(\d+)\s+.*model\s+(\d+) cycles", line)
        if m:
            model[int(m.group(1))] = int(m.group(2))
    meas = {}
    for line in open(path, errors="replace"):
        m = re.match(r"(.+?)\s+([0-9.]+) cycles per dependent op", line)
        if not m:
            continue
        for lab, case in LABEL.items():
            if m.group(1).strip().startswith(lab):
                meas[case] = float(m.group(2))
    if not meas:
        sys.exit("no warp_latency lines found in %s" % path)
    print("case  measured  model  ratio")
    for case in sorted(meas):
        if case in model:
            print("%4d  %8.1f  %5d  %5.2f" % (case, meas[case], model[case], meas[case] / model[case]))
    print("\nextra cycles per TAKEN branch beyond the model (instruction fetch / redirect), by loop footprint:")
    for case, (nb, fp) in BRANCHES.items():
        if case in meas and case in model:
            print("  footprint %-7s %6.1f cycles per branch" % (fp, (meas[case] - model[case]) / nb))
    if 112 in meas and 110 in meas:
        print("\nstraight-line code: %.2f cycles per instruction at 1 KB, %.2f at 32 KB" % (meas[110] / 64.0, meas[112] / 2048.0))


if __name__ == "__main__":
    main()
