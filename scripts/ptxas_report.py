#!/usr/bin/env python
"""Per-kernel resource usage (`ptxas -v`) of libsttode_sm100.so -> profiles/rNN_ptxas_resources.md (no GPU needed).

    python scripts/ptxas_report.py [round]

Compiles the library's sources with the release flags + `-Xptxas -v` into a scratch file (the in-tree .so is not touched) and tabulates
registers per thread, static shared memory, stack frame and spill bytes of every kernel."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sttode_b200 import build as B      # noqa: E402


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "02"
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [B._nvcc()] + B.NVCC_FLAGS + ["-Xptxas", "-v", "-o", os.path.join(tmp, "x.so")] + [os.path.join(B.CSRC, s) for s in B.SOURCES]
        log = subprocess.run(cmd, capture_output=True, text=True, check=True).stderr
    rows, cur = [], None
    for ln in log.splitlines():
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", ln)
        if m:
            cur = {"name": m.group(1), "regs": 0, "smem": 0, "stack": 0, "st": 0, "ld": 0, "bars": 0}
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m:
            cur["stack"], cur["st"], cur["ld"] = map(int, m.groups())
        m = re.search(r"Used (\d+) registers, used (\d+) barriers", ln)
        if m:
            cur["regs"], cur["bars"] = int(m.group(1)), int(m.group(2))
            s = re.search(r"(\d+) bytes smem", ln)
            cur["smem"] = int(s.group(1)) if s else 0
    names = subprocess.run(["c++filt", "-p"], input="\n".join(r["name"] for r in rows), capture_output=True, text=True, check=True).stdout.splitlines()
    for r, n in zip(rows, names):
        r["name"] = n.replace("sttode::", "")
    rows.sort(key=lambda r: (-r["regs"], r["name"]))
    lines = [f"# ptxas resource usage of libsttode_sm100.so (round {rnd})", "",
             "`nvcc " + " ".join(B.NVCC_FLAGS) + " -Xptxas -v`, one row per kernel.  Static shared memory only: the tensor-core kernels",
             "and the row-tile kernels request their (dynamic) shared memory at launch (DESIGN.md §4).  Regenerate with `python scripts/ptxas_report.py`.", "",
             "| kernel | registers / thread | named barriers | static smem (B) | stack frame (B) | spill stores (B) | spill loads (B) |",
             "|---|---|---|---|---|---|---|"]
    for r in rows:
        lines.append(f"| `{r['name']}` | {r['regs']} | {r['bars']} | {r['smem']} | {r['stack']} | {r['st']} | {r['ld']} |")
    path = os.path.join(ROOT, "profiles", f"r{rnd}_ptxas_resources.md")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(path, len(rows), "kernels")


if __name__ == "__main__":
    main()
