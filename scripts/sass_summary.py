#!/usr/bin/env python
"""Static SASS counts per kernel of libaisp_b200.so (no GPU needed):
    python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "adaptiveisp_b200", "csrc", "libaisp_b200.so")
COLS = ["UTMALDG", "UTMAPF", "UBLKPF", "SYNCS", "LDGSTS", "FADD2", "FMUL2", "FFMA2", "MUFU", "SHFL", "LDG.E.128", "STG.E.128",
        "LDS.64", "LDS.128", "STL", "LDL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    names = iter(filt)
    agg = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            full = next(names)
            short = re.sub(r"^void ", "", full)
            short = re.sub(r"\(.*", "", short).replace("aisp::", "")
            cur = agg.setdefault(short, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["instrs"] += 1
            for c in COLS:
                if c in ("LDG.E.128", "STG.E.128"):
                    hit = op.startswith(c[:3]) and ".128" in op
                elif c in ("LDS.64", "LDS.128"):
                    hit = op.startswith("LDS") and op.endswith(c[3:])
                else:
                    hit = op == c or op.startswith(c + ".")
                if hit:
                    cur[c] += 1
    print("SASS evidence, libaisp_b200.so (sm_100a), round 2 final -- static instruction counts per kernel (cuobjdump -sass;")
    print("regenerate with scripts/sass_summary.py).  UTMALDG / UTMAPF = TMA tensor load / TMA L2 prefetch; UBLKPF = bulk L2 prefetch;")
    print("SYNCS = mbarrier; LDGSTS = cp.async; FADD2 / FMUL2 / FFMA2 = Blackwell packed fp32; STL / LDL = local-memory spills.")
    print("No HMMA / UTC*MMA anywhere: no stage is a dense contraction.\n")
    print(f"{'kernel':78s} {'instrs':>7s} " + " ".join(f"{c:>9s}" for c in COLS))
    for k, c in sorted(agg.items()):
        print(f"{k[:78]:78s} {c['instrs']:7d} " + " ".join(f"{c[x]:9d}" for x in COLS))


if __name__ == "__main__":
    main()
