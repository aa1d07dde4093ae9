#!/usr/bin/env python3
"""tools/cmp_sass.py OLD.o NEW.o -- are the kernels of two builds of csrc/qlfc.cu the same machine code?
Used at the end of round 1 (no GPU minutes left) to show that the GPU-verified kernels were not touched by later host-side
and experimental additions: every kernel of the verified commit must have byte-identical SASS text in the new object
(the static encoder became a template on the way: q_encode5 is matched with q_encode5<LayoutFull, false>)."""
import re, subprocess, sys
def funcs(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m: cur = m.group(1); d[cur] = []; continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);\s*/\*", line)
        if m and cur: d[cur].append(re.sub(r"\s+", " ", m.group(1)))
    return d
a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
def key(n):   # strip the per-TU hash of the anonymous namespace and template spelling differences
    n = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_qlfc_cu_[0-9a-f]+", "NS", n)
    return n
ka = {key(k): k for k in a}; kb = {key(k): k for k in b}
for k in sorted(ka):
    # the encoder became a template: q_encode5 -> q_encode5<false>
    cand = [k2 for k2 in kb if k2 == k or (("q_encode5" in k) and ("q_encode5" in k2) and ("Lj5ELj5ELi12" in k2 or "q_encode5ILb0" in k2) and "Lb0" in k2)]
    if not cand: print("MISSING in new:", k); continue
    same = a[ka[k]] == b[kb[cand[0]]]
    print(("SAME  " if same else "DIFF  ") + "%5d instr  " % len(a[ka[k]]) + k[:110])
