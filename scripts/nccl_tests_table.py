#!/usr/bin/env python
"""Turns the nccl-tests logs of scripts/run_nccl_tests.sh into one markdown table per collective:
out-of-place time / bus bandwidth of system NCCL next to the uccl_b200 drop-in, and the `#wrong` column."""
import os
import re
import sys


def parse(path):
    rows, errs = {}, None
    if not os.path.exists(path):
        return rows, "missing"
    for ln in open(path):
        p = ln.split()
        # size count type redop root  time algbw busbw #wrong  time algbw busbw #wrong
        if len(p) >= 12 and p[0].isdigit() and p[1].isdigit():
            try:
                rows[int(p[0])] = dict(t=float(p[-8]), bus=float(p[-6]), wrong=p[-5], t_ip=float(p[-4]), bus_ip=float(p[-2]), wrong_ip=p[-1])
            except ValueError:
                pass
        m = re.search(r"Out of bounds values\s*:\s*(\d+)\s*(\w+)", ln)
        if m:
            errs = f"{m.group(1)} {m.group(2)}"
    return rows, errs


def main():
    d = sys.argv[1]
    tests = sorted({f[len("nccl_"):-4] for f in os.listdir(d) if f.startswith("nccl_") and f.endswith(".txt")})
    for t in tests:
        a, ea = parse(os.path.join(d, f"nccl_{t}.txt"))
        b, eb = parse(os.path.join(d, f"uccl_b200_{t}.txt"))
        print(f"### {t}_perf (out-of-place; in-place in parentheses)  -- out-of-bounds: NCCL {ea}, uccl_b200 {eb}\n")
        print("| bytes | NCCL us | NCCL busbw GB/s | uccl_b200 us | uccl_b200 busbw GB/s | speed-up | #wrong ours |")
        print("|---:|---:|---:|---:|---:|---:|---:|")
        for sz in sorted(set(a) | set(b)):
            x, y = a.get(sz), b.get(sz)
            if not x or not y:
                continue
            print(f"| {sz} | {x['t']:.1f} ({x['t_ip']:.1f}) | {x['bus']:.1f} ({x['bus_ip']:.1f}) | {y['t']:.1f} ({y['t_ip']:.1f}) | "
                  f"{y['bus']:.1f} ({y['bus_ip']:.1f}) | {x['t'] / y['t']:.2f}x ({x['t_ip'] / y['t_ip']:.2f}x) | {y['wrong']} ({y['wrong_ip']}) |")
        print()


if __name__ == "__main__":
    main()
