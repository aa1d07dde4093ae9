#!/usr/bin/env python3
"""tools/sass_t1w.py OBJ KERNEL-SUBSTRING [--from ADDR --to ADDR] [--bb] [--trace A,B,C...]

Static single-warp timing of a SASS region, after the event-step model of /opt/skills/guides/B300_MICROARCH.md
("Single-warp T_1w"): every instruction carries its own stall count and scoreboard fields in the control bits of the
128-bit encoding (bits 105-108 stall, 109 yield, 110-112 write barrier, 113-115 read barrier, 116-121 wait mask), so the
time one LONE warp -- the situation of a QLFC coder warp -- needs for a straight-line path is

    T = max(T + stall, scoreboards it waits for);  a variable-latency producer arms its barrier at T + latency(op).

Used without a GPU to compare coder variants: per basic block the instruction count, the sum of the stalls (= issue time if
no scoreboard ever binds) and the modelled time with scoreboards, and for a given chain of basic blocks (--trace, block start
addresses in execution order, scoreboard state carried across) the modelled cycles of that path.
Latencies of the variable-latency classes are the B200 measurements of tools/warp_latency.cu (profiles/r1f_warp_latency.txt)
minus loop overhead; they are parameters of the model, not facts about a particular kernel.
"""
import re
import subprocess
import sys

LAT = {"LDS": 26, "LDG": 400, "LD": 400, "LDC": 30, "LDCU": 30, "SHFL": 26, "MATCH": 40, "VOTE": 12, "VOTEU": 12, "REDUX": 22, "STS": 6, "STG": 6, "ST": 6,
       "POPC": 14, "FLO": 14, "BREV": 14, "MUFU": 18, "S2R": 20, "S2UR": 20, "CS2R": 6, "I2F": 14, "F2I": 14, "ATOMS": 40, "ATOMG": 500, "ATOM": 500, "RED": 6,
       "R2UR": 12, "BAR": 20, "WARPSYNC": 6, "NANOSLEEP": 50, "LDSM": 30, "BMSK": 6, "IMAD.WIDE": 6, "SYNCS": 30, "PRMT": 6, "CCTL": 50, "ERRBAR": 20, "MEMBAR": 100,
       "CALL": 6, "RET": 6, "BMOV": 12, "DEPBAR": 0}
RBAR_LAT = 6          # operand-read latch window of a variable-latency op (guide: "T + 6")


class Ins:
    __slots__ = ("addr", "text", "op", "pred", "stall", "yld", "wbar", "rbar", "wait", "target")


def disasm(obj, want):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs, cur, name = {}, None, None
    pend = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            cur = funcs.setdefault(name, [])
            pend = None
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", line)
        if m and cur is not None:
            i = Ins()
            i.addr, i.text = int(m.group(1), 16), re.sub(r"\s+", " ", m.group(2)).strip()
            pend = i
            continue
        m = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", line)
        if m and pend is not None:
            hi = int(m.group(1), 16)
            i = pend
            pend = None
            i.stall, i.yld = (hi >> 41) & 15, (hi >> 45) & 1
            i.wbar, i.rbar, i.wait = (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 63
            t = i.text
            pm = re.match(r"(@!?U?P\d+)\s+(.*)", t)
            i.pred = pm.group(1) if pm else None
            body = pm.group(2) if pm else t
            i.op = body.split()[0]
            tm = re.search(r"\b(?:BRA|BSSY|BRX|CALL\.REL\.NOINC|BRA\.U|BRA\.DIV)\S*\s+(?:.*?,\s*)?(?:`\(\S+\)|0x([0-9a-f]+))", body)
            i.target = int(tm.group(1), 16) if (tm and tm.group(1) and i.op.startswith("BRA")) else None
            cur.append(i)
    hits = [n for n in funcs if want in n]
    if len(hits) != 1:
        sys.exit("kernel substring %r matches %d functions:\n  %s" % (want, len(hits), "\n  ".join(hits[:20])))
    return hits[0], funcs[hits[0]]


def lat_of(op):
    base = op.split(".")[0]
    if op.startswith("IMAD.WIDE"):
        return 6
    return LAT.get(base, 12)


def blocks(ins):
    starts = {ins[0].addr}
    for k, i in enumerate(ins):
        if i.target is not None:
            starts.add(i.target)
        if i.op.split(".")[0] in ("BRA", "EXIT", "RET", "BRX", "BSYNC", "BREAK", "CALL") and k + 1 < len(ins):
            starts.add(ins[k + 1].addr)
    order = sorted(starts)
    idx = {a: n for n, a in enumerate(order)}
    bbs = [[] for _ in order]
    cur = 0
    for i in ins:
        if i.addr in idx:
            cur = idx[i.addr]
        bbs[cur].append(i)
    return [b for b in bbs if b]


def simulate(seq, T=0, sb=None, verbose=False):
    """seq: instructions in execution order.  Returns (T, sb, exposed) -- exposed = cycles where a scoreboard wait was binding."""
    sb = dict(sb or {})
    exposed = 0
    for i in seq:
        arm = max([sb.get(s, 0) for s in range(6) if (i.wait >> s) & 1] or [0])
        t_issue = max(T, arm)
        if arm > T:
            exposed += arm - T
            if verbose:
                print("      wait %3d cycles at %04x  %s" % (arm - T, i.addr, i.text[:70]))
        if i.wbar < 6:
            sb[i.wbar] = max(sb.get(i.wbar, 0), t_issue + lat_of(i.op))
        if i.rbar < 6:
            sb[i.rbar] = max(sb.get(i.rbar, 0), t_issue + RBAR_LAT)
        T = t_issue + max(i.stall, 1)
    return T, sb, exposed


def main():
    a = sys.argv[1:]
    if len(a) < 2:
        sys.exit(__doc__)
    name, ins = disasm(a[0], a[1])
    lo = int(a[a.index("--from") + 1], 16) if "--from" in a else 0
    hi = int(a[a.index("--to") + 1], 16) if "--to" in a else 1 << 30
    print("# %s: %d instructions" % (name[:100], len(ins)))
    bbs = blocks(ins)
    by_start = {b[0].addr: b for b in bbs}
    if "--trace" in a:
        T, sb, exp_total = 0, {}, 0
        n = 0
        for tok in a[a.index("--trace") + 1].split(","):
            rep = 1
            if "*" in tok:
                tok, r = tok.split("*")
                rep = int(r)
            b = by_start[int(tok, 16)]
            for _ in range(rep):
                T, sb, e = simulate(b, T, {k: v for k, v in sb.items()}, verbose="-v" in a)
                exp_total += e
                n += len(b)
        print("trace: %d instructions, %d cycles modelled (%d exposed to scoreboards) = %.2f cycles / instruction" % (n, T, exp_total, T / max(n, 1)))
        return
    if "--bb" in a:
        print("# start  end   instr  sum(stall)  modelled  exposed  last instruction")
        for b in bbs:
            if b[0].addr < lo or b[0].addr > hi:
                continue
            T, _, e = simulate(b)
            print("  %04x  %04x  %5d  %9d  %8d  %7d  %s" % (b[0].addr, b[-1].addr, len(b), sum(max(i.stall, 1) for i in b), T, e, b[-1].text[:60]))
        return
    for i in ins:
        if lo <= i.addr <= hi:
            print("%04x  s%-2d %s w%s r%s wait=%02x  %s" % (i.addr, i.stall, "Y" if i.yld else " ", i.wbar if i.wbar < 6 else "-", i.rbar if i.rbar < 6 else "-", i.wait, i.text))


if __name__ == "__main__":
    main()
