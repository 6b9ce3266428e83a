"""Zero-one check of the lane-local sorting networks in csrc/sort_net.cuh (presort_lane): parses the AVR_CE lists out
of the header and runs each over all 2^n zero-one inputs (a comparator network sorts every input iff it sorts those).

    python tools/check_sort_nets.py
"""
import os
import re
import sys

import numpy as np

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "adaptive-volume-rendering_b200", "csrc", "sort_net.cuh")


def networks():
    text = open(HDR).read()
    body = text[text.index("presort_lane"):text.index("#undef AVR_CE")]
    for m in re.finditer(r"EPF == (\d+)\) \{(.*?)return true;", body, re.S):
        yield int(m.group(1)), [(int(a), int(b)) for a, b in re.findall(r"AVR_CE\((\d+), (\d+)\)", m.group(2))]


def sorts(n, net):
    x = np.arange(1 << n, dtype=np.uint32)
    bits = [(x >> i) & 1 for i in range(n)]
    for a, b in net:
        bits[a], bits[b] = bits[a] & bits[b], bits[a] | bits[b]
    return all(bool((bits[i] <= bits[i + 1]).all()) for i in range(n - 1))


def main():
    found = list(networks())
    assert [n for n, _ in found] == [4, 8, 16], found
    for n, net in found:
        ok = sorts(n, net) and all(a < b < n for a, b in net)
        print(f"{n} keys: {len(net)} compare-exchanges, sorts every zero-one input: {ok}")
        if not ok:
            sys.exit(1)


if __name__ == "__main__":
    main()
