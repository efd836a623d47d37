#!/usr/bin/env python
"""tools/sass_mix.py -- static instruction mix of a kernel's loops from `cuobjdump -sass` (no GPU needed).

    python tools/sass_mix.py seal_b200/csrc/sb_engine.o '<mangled kernel name>'

Every hot kernel of this path is bound by the issue rate of the integer-multiply pipe, so the number of
IMAD-family instructions per loop iteration is a direct predictor of its run time; this prints that count and the rest of
the mix for every loop (backward branch) of the kernel."""
import collections
import re
import subprocess
import sys


def main(obj, fun):
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
    lines = [l for l in sass.splitlines() if re.search(r"/\*[0-9a-f]{4,5}\*/", l)]
    addr = lambda l: int(re.search(r"/\*([0-9a-f]{4,5})\*/", l).group(1), 16)

    def op(l):
        m = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if not m:
            return None
        p = m.group(1).split(".")
        return "IMAD." + p[1] if p[0] == "IMAD" and len(p) > 1 and p[1] in ("WIDE", "X", "MOV", "IADD", "SHL", "HI") else p[0]

    print(f"{fun}: {len(lines)} instructions")
    loops = []
    for l in lines:
        if "BRA" in l:
            t = re.search(r"0x([0-9a-f]+)", l.split("BRA")[1])
            if t and int(t.group(1), 16) < addr(l):
                loops.append((int(t.group(1), 16), addr(l)))
    for a, b in sorted(loops, key=lambda x: x[0] - x[1])[:4]:
        c = collections.Counter(o for o in (op(l) for l in lines if a <= addr(l) <= b) if o)
        fam = sum(v for k, v in c.items() if k.startswith("IMAD"))
        over = c["IMAD.X"] + c["IMAD.MOV"] + c["IMAD.IADD"] + c["IMAD.SHL"]
        print(f"loop {a:#x}..{b:#x}: {sum(c.values())} instructions, IMAD family {fam} (WIDE {c['IMAD.WIDE']}, IMAD {c['IMAD']}, "
              f"HI {c['IMAD.HI']}, adds/moves on the multiply pipe {over}), IADD3 {c['IADD3']}, LDG {c['LDG']}, LDS {c['LDS']}, STS {c['STS']}")
        print("   " + ", ".join(f"{k} {v}" for k, v in c.most_common(14)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
