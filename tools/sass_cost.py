#!/usr/bin/env python
"""tools/sass_cost.py -- static cost estimate of a kernel's loops from `cuobjdump -sass` (no GPU needed).

    python tools/sass_cost.py <object or .so> '<substring of the mangled kernel name>' [butterflies per iteration]

Cost model measured with tools/issue_microbench.cu on B200 (profiles/r02_issue_microbench.md): per warp, IMAD.WIDE / IMAD.HI
occupy the integer-multiply pipe for 4 clocks, every other IMAD-family instruction (IMAD, IMAD.X, IMAD.MOV, IMAD.IADD, IMAD.SHL)
for 2; instructions of the other pipes (integer ALU, load/store, branches) overlap with it only partly and add ~0.9 clocks each.
    estimate = 4*WIDE + 4*HI + 2*IMAD_other + 0.9*everything_else      (clocks per warp per loop iteration per SM sub-partition)
"""
import collections
import re
import subprocess
import sys


def kernels(obj):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, body = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and cur:
            body[cur].append((int(m.group(1), 16), m.group(2)))
    return body


def classify(text):
    t = text.split()
    op = t[1] if t[0].startswith("@") else t[0]
    if op.startswith("IMAD.WIDE"):
        return "IMAD.WIDE"
    if op.startswith("IMAD.HI"):
        return "IMAD.HI"
    p = op.split(".")
    if p[0] == "IMAD":
        return "IMAD." + p[1] if len(p) > 1 and p[1] in ("X", "MOV", "IADD", "SHL") else "IMAD"
    return p[0]


def cost(c):
    fma = 4 * c["IMAD.WIDE"] + 4 * c["IMAD.HI"] + 2 * sum(v for k, v in c.items() if k.startswith("IMAD") and k not in ("IMAD.WIDE", "IMAD.HI"))
    other = sum(v for k, v in c.items() if not k.startswith("IMAD"))
    return fma, other, fma + 0.9 * other


def main(obj, pat, per=None):
    for fn, ins in kernels(obj).items():
        if pat not in fn:
            continue
        loops = []
        for addr, t in ins:
            m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < addr:
                loops.append((int(m.group(1), 16), addr))
        call = collections.Counter(classify(t) for a, t in ins)
        fma, other, est = cost(call)
        print(f"{fn}: {len(ins)} instructions, {len(loops)} loops; whole kernel: multiply-pipe clocks {fma}, other instr {other}, estimate {est:.0f}"
              + (f" = {est / float(per):.1f} per unit" if per else ""))
        for lo, hi in sorted(loops, key=lambda l: l[0] - l[1])[:3]:
            c = collections.Counter(classify(t) for a, t in ins if lo <= a <= hi)
            fma, other, est = cost(c)
            line = f"  loop {lo:#x}..{hi:#x}: {sum(c.values())} instr; multiply-pipe clocks {fma}, other instr {other}, estimate {est:.0f} clk"
            if per:
                line += f" = {est / float(per):.1f} per unit"
            print(line)
            print("     " + ", ".join(f"{k} {v}" for k, v in c.most_common(16)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
