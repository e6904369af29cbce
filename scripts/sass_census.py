#!/usr/bin/env python
"""Per-kernel SASS instruction census of libsttode_sm100.so -> profiles/rNN_sass_census.md (cuobjdump -sass; no GPU needed).

    python scripts/sass_census.py [round]

Counts the mnemonics that show what a kernel runs on: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit),
UBLKCP (cp.async.bulk 1-D), UTMALDG (tensor-map TMA), SYNCS (mbarrier), FFMA / HFMA2 / MUFU, LDG/STG, RED/ATOM."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sttode_b200", "lib", "libsttode_sm100.so")
COLS = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "FFMA", "HFMA2", "MUFU", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR"]


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "02"
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = counts.setdefault(name.split("(")[0], collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for c in COLS:
                if op == c or op.startswith(c + "."):
                    cur[c] += 1
    lines = [f"# SASS census of libsttode_sm100.so (round {rnd})", "",
             "`cuobjdump -sass sttode_b200/lib/libsttode_sm100.so`, instruction counts per kernel (static, not executed counts).",
             "UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D bulk copy), UTMALDG = tensor-map TMA,",
             "SYNCS = mbarrier ops.  Regenerate with `python scripts/sass_census.py`.", "",
             "| kernel | total | " + " | ".join(COLS) + " |", "|---|---|" + "---|" * len(COLS)]
    for name, c in sorted(counts.items(), key=lambda kv: -kv[1]["UTCHMMA"] * 100000 - kv[1]["total"]):
        short = name.replace("sttode::", "")
        lines.append(f"| `{short}` | {c['total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in COLS) + " |")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    lines.append(f"| **all {len(counts)} kernels** | {tot['total']} | " + " | ".join(str(tot[k]) for k in COLS) + " |")
    path = os.path.join(ROOT, "profiles", f"r{rnd}_sass_census.md")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(path)


if __name__ == "__main__":
    main()
