"""Executed warp instructions and stall samples per SOURCE LINE of one kernel, from an `ncu --set full
--import-source on` report joined with the object's line table.

    python tools/ncu_lines.py <report.ncu-rep> <object.o> <substring of the mangled kernel name> [--min-pct 1.0] [--per N]

The report's source page lists the kernel's SASS in address order with `Instructions Executed` and the
stall samples of each instruction; `nvdisasm -g` of the object (built with -lineinfo from the same
sources) gives the source line of each instruction.  The two listings are joined by position (and checked
opcode by opcode).  `--per N` divides the counts by N (e.g. the number of rays or samples of the launch).  `lanes` is the average
number of active threads of the line's instructions: well under 32 on a hot line means divergence (that is how the
packed coarse sampler's twice-executed block was found, profiles/r02_packed_samplers_pass3.md).
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_lines import disassemble, kernel_rows  # noqa: E402


def report_rows(rep, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *extra], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    res = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        d = dict(zip(hdr, r))
        m = re.match(r"\s*(@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", d["Source"])
        res.append((m.group(2) if m else "?", float(d["Instructions Executed"] or 0), float(d["# Samples"] or 0),
                    float(d.get("Thread Instructions Executed") or 0)))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("--min-pct", type=float, default=1.0)
    ap.add_argument("--per", type=float, default=0.0)
    ap.add_argument("--ncu", default="", help='filters for a report with several launches, e.g. "--kernel-name regex:grp --launch-skip 3 --launch-count 1"')
    a = ap.parse_args()
    name, rows = kernel_rows(disassemble(a.obj), a.kernel)
    rep = report_rows(a.rep, a.ncu.split())
    if len(rows) != len(rep):
        print(f"warning: object has {len(rows)} instructions, report {len(rep)} (different builds?)", file=sys.stderr)
    n = min(len(rows), len(rep))
    bad = sum(1 for i in range(n) if rows[i][1] != rep[i][0])
    if bad:
        print(f"warning: {bad} opcode mismatches in the join", file=sys.stderr)
    inst = collections.Counter()
    thr = collections.Counter()
    samp = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    for i in range(n):
        src = rows[i][2]
        inst[src] += rep[i][1]
        samp[src] += rep[i][2]
        thr[src] += rep[i][3]
        ops[src][rows[i][1]] += rep[i][1]
    ti, ts = sum(inst.values()), sum(samp.values())
    unit = a.per if a.per else 1.0
    print(f"{name}\nexecuted warp instructions {ti:.4g}" + (f" = {ti / unit:.2f} per unit" if a.per else "") + f", stall samples {ts:.0f}")
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "adaptive-volume-rendering_b200", "csrc")
    cache = {}
    for src in sorted(inst):
        pi, ps = 100 * inst[src] / max(ti, 1), 100 * samp[src] / max(ts, 1)
        if pi < a.min_pct and ps < a.min_pct:
            continue
        f, l = src
        p = os.path.join(root, f)
        if f not in cache:
            cache[f] = open(p).read().split("\n") if os.path.exists(p) else None
        text = cache[f][l - 1].strip()[:70] if cache[f] and l <= len(cache[f]) else ""
        top = ", ".join(f"{o} {c / unit:.3g}" for o, c in ops[src].most_common(3))
        lanes = thr[src] / inst[src] if inst[src] else 0.0
        print(f"{f}:{l:<4} inst {pi:5.1f}%  stalls {ps:5.1f}%  lanes {lanes:4.1f}  [{top}]  {text}")


if __name__ == "__main__":
    main()
