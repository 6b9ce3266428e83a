"""Static SASS accounting for one kernel: instructions per source line inside its main loop.

    python tools/sass_lines.py <object.o> <substring of the mangled kernel name> [--min 6] [--all]

Uses `cuobjdump -xelf` + `nvdisasm -g` (the objects are built with -lineinfo).  The "main loop" is
the range of the longest backward branch that lies before the kernel's cold section (ptxas puts
the slow paths of full-mask shuffles and other rarely taken code after the first unconditional
EXIT).  This is what the per-kernel instruction budgets in profiles/ and DESIGN.md were read from;
it needs no GPU.
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile


def disassemble(obj):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    return subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], check=True, capture_output=True, text=True).stdout


def kernel_rows(text, want):
    rows, cur, active, name = [], None, False, None
    for line in text.split("\n"):
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            active = want in m.group(1) and name in (None, m.group(1))
            if active:
                name = m.group(1)
            cur = None
            continue
        if not active:
            continue
        m = re.search(r'//## File "(.*?)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P[0-9T]\s+)?([A-Z0-9_]+)(.*)", line)
        if m and cur:
            rows.append((int(m.group(1), 16), m.group(3), cur, m.group(4), bool(m.group(2))))
    return name, rows


def main_loop(rows):
    exits = [i for i, r in enumerate(rows) if r[1] == "EXIT" and not r[4]]
    hot_end = rows[exits[0]][0] if exits else rows[-1][0]
    best = None
    for a, op, _, rest, _p in rows:
        m = re.search(r"(0x[0-9a-f]+)", rest)
        if op == "BRA" and m and a <= hot_end:
            t = int(m.group(1), 16)
            if t < a and (best is None or a - t > best[1] - best[0]):
                best = (t, a)
    return best if best else (rows[0][0], hot_end)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("--min", type=int, default=6)
    ap.add_argument("--all", action="store_true", help="whole kernel instead of the main loop")
    a = ap.parse_args()
    name, rows = kernel_rows(disassemble(a.obj), a.kernel)
    if not rows:
        sys.exit(f"no kernel matching {a.kernel!r}")
    lo, hi = (rows[0][0], rows[-1][0]) if a.all else main_loop(rows)
    body = [r for r in rows if lo <= r[0] <= hi]
    ops = collections.Counter(r[1] for r in body)
    print(f"{name}\n{len(rows)} instructions, range {lo:#x}..{hi:#x}: {len(body)}")
    print("opcodes:", ", ".join(f"{o} {c}" for o, c in ops.most_common(16)))
    per = collections.defaultdict(collections.Counter)
    for _, op, src, _, _p in body:
        per[src][op] += 1
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "adaptive-volume-rendering_b200", "csrc")
    cache = {}
    for (f, l), c in sorted(per.items()):
        n = sum(c.values())
        if n < a.min:
            continue
        p = os.path.join(root, f)
        if f not in cache:
            cache[f] = open(p).read().split("\n") if os.path.exists(p) else None
        text = cache[f][l - 1].strip()[:78] if cache[f] and l <= len(cache[f]) else ""
        print(f"{f}:{l:<4} {n:>4}  {dict(c.most_common(3))}  {text}")


if __name__ == "__main__":
    main()
