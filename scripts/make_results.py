#!/usr/bin/env python
"""Regenerate profiles/RESULTS.md from the JSON files benchmarks write (profiles/*.json).

    python scripts/make_results.py            # reads profiles/, writes profiles/RESULTS.md
"""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(name):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return None
    txt = open(path).read().strip()
    try:
        return json.loads(txt)
    except json.JSONDecodeError:  # bench.py prints one JSON line, possibly after launcher chatter
        for line in reversed(txt.splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return None


def us(v):
    return f"{v:.1f} µs"


def ar_table(d):
    out = [f"## AllReduce, {d.get('dtype', 'bf16')} sum, {d['n_gpus']}×B200 (NVLS multicast = {d.get('nvls')})",
           "| bytes | ours sym (auto) | algo | ours plain (staged/LL best) | NCCL | speed-up sym | busbw ours / NCCL (GB/s) |",
           "|---:|---:|---|---:|---:|---:|---:|"]
    for r in d["rows"]:
        plain = [r[k]["us"] for k in ("staged_p2p", "staged_nvls", "oneshot_ll", "oneshot_mc") if k in r]
        a, n = r["auto_sym"], r["nccl"]
        out.append(f"| {r['bytes']} | {us(a['us'])} | {a.get('algo', '')} | {us(min(plain)) if plain else '-'} | {us(n['us'])} | "
                   f"{n['us'] / a['us']:.2f}× | {a['busbw']:.1f} / {n['busbw']:.1f} |")
    return out


def coll_table(d):
    rows = d["rows"]
    cols = [k for k in ("sym_out", "sym_in", "plain", "plain_noll", "plain_ll", "nccl") if any(k in r for r in rows)]
    out = [f"## {d['coll']}, {d.get('dtype', 'bf16')}, {d['n_gpus']}×B200 (total bytes; µs and busbw GB/s)", "",
           "| bytes | " + " | ".join(cols) + " |", "|---:|" + "---:|" * len(cols)]
    for r in rows:
        cells = [f"{r[c]['us']:.1f} µs ({r[c]['busbw']:.0f})" if c in r else "-" for c in cols]
        out.append(f"| {r['bytes']} | " + " | ".join(cells) + " |")
    if "plain_noll" in cols:
        out += ["", "`plain_noll` / `plain_ll`: the same call with the LL-packet exchange forced off / on "
                "(`plain` = built-in threshold)."]
    return out


def ep_table(d):
    out = [f"## EP dispatch / combine, {d['n_gpus']}×B200: {d['tokens']} tokens/rank, hidden {d['hidden']}, "
           f"top-{d['topk']}, {d['experts']} experts", "",
           "| CTAs | dispatch bf16→fp8 fused | dispatch bf16 | combine bf16 | recv tokens |", "|---:|---:|---:|---:|---:|"]
    by = {}
    lls = [r for r in d["rows"] if r.get("ll")]
    for r in d["rows"]:
        if not r.get("ll"):
            by.setdefault(r["sms"], {})[r["mode"]] = r
    for sms, m in sorted(by.items()):
        f8, b = m.get("fp8_fused"), m.get("bf16")
        c1 = f"{f8['dispatch_us']:.0f} µs ({f8['dispatch_GBps']:.0f} GB/s)" if f8 else "-"
        c2 = f"{b['dispatch_us']:.0f} µs ({b['dispatch_GBps']:.0f} GB/s)" if b else "-"
        c3 = f"{b['combine_us']:.0f} µs ({b['combine_GBps']:.0f} GB/s)" if b and "combine_us" in b else "-"
        out.append(f"| {sms} | {c1} | {c2} | {c3} | {(b or f8)['num_recv']} |")
    for ll in lls:
        out += ["", f"Low-latency, {ll.get('tokens', 128)} tokens/rank, fp8={ll.get('use_fp8')}: "
                f"dispatch {ll['dispatch_us']:.1f} µs, combine {ll['combine_us']:.1f} µs"]
    return out


def ctas_table(d):
    algos = ("twoshot_nvls", "twoshot_p2p", "staged_nvls")
    counts = sorted({int(k.split("@")[1]) for r in d["rows"] for k in r if "@" in k})
    out = [f"### CTA-count sensitivity at {d['n_gpus']} GPUs (µs)", "",
           "| bytes | " + " | ".join(f"{a} @{'/'.join(map(str, counts))}" for a in algos) + " | NCCL |",
           "|---:|" + "---|" * len(algos) + "---:|"]
    for r in d["rows"]:
        cells = [" / ".join(f"{r[f'{a}@{c}']['us']:.0f}" if f"{a}@{c}" in r else "-" for c in counts) for a in algos]
        out.append(f"| {r['bytes']} | " + " | ".join(cells) + f" | {r['nccl']['us']:.0f} |")
    return out


def p2p_table(d):
    out = [f"## P2P engine GPU{d['devices'][0]}→GPU{d['devices'][1]} (host-timed per transfer incl. completion poll)", "",
           "| bytes | blocks | write (TMA kernel) | read (TMA kernel) | cudaMemcpyPeer |", "|---:|---:|---:|---:|---:|"]
    for r in d["rows"]:
        out.append(f"| {r['bytes']} | {r.get('blocks', 1)} | {r['write']['GBps']:.0f} GB/s | {r['read']['GBps']:.0f} GB/s | "
                   f"{r['memcpy_peer']['GBps']:.0f} GB/s |")
    return out


def bench_line(d):
    c = d.get("clocks", {})
    e = d.get("e2e", {})
    mm = d.get("step_us_min_med_max")
    spread = f" (min/med/max {mm[0]:.0f}/{mm[1]:.0f}/{mm[2]:.0f})" if mm else ""
    return (f"| {d['n_gpus']} | {d['value'] / 1e6:.1f} M tok/s | {d['ms_per_step'] * 1e3:.1f} µs{spread} | "
            f"{d.get('dispatch_us', 0):.1f} / {d.get('combine_us', 0):.1f} µs | {('%.2f' % d['vs_baseline']) if d.get('vs_baseline') else '-'} | "
            f"{e.get('value', 0) / 1e6:.1f} M tok/s | {d.get('gpu_launches')} | {c.get('sm_mhz')} MHz {c.get('reasons')} |")


def main():
    out = ["# RESULTS — measured on B200 (sm_100a), NCCL 2.28.9 as the baseline on the same box",
           "Device-timed (CUDA events), max over ranks, after warm-up; bus bandwidth uses the nccl-tests",
           "formulas. `sym` = buffers from the symmetric heap (`Communicator.empty` / `ncclMemAlloc` /",
           "`comm.use_mem_pool()`), `plain` = ordinary `cudaMalloc`/torch tensors (staged or LL-packet kernels).",
           "Generated by `scripts/make_results.py` from the JSON files in this directory.", ""]
    benches = [load(f"bench{n}.json") for n in (1, 2, 4, 8)]
    benches = [b for b in benches if b]
    if benches:
        out += ["## bench.py (EP dispatch + combine, DeepSeek-V3 shape; reference 8×B200: 571 + 727 µs)", "",
                "| GPUs | tokens/s | step | dispatch / combine | vs published baseline | end-to-end | launches/region | clocks |",
                "|---:|---:|---:|---:|---:|---:|---:|---|"]
        out += [bench_line(b) for b in benches] + [""]
    for n in (8, 4, 2):
        d = load(f"ar{n}.json")
        if d:
            out += ar_table(d) + [""]
        d = load(f"ar{n}_ctas.json")
        if d:
            out += ctas_table(d) + [""]
    for coll in ("allgather", "reduce_scatter", "alltoall", "broadcast"):
        for n in (8, 4, 2):
            d = load(f"{coll}{n}.json")
            if d:
                out += coll_table(d) + [""]
    for n in (8, 4, 2, 1):
        d = load(f"ep{n}.json")
        if d:
            out += ep_table(d) + [""]
    for n in (8, 4, 2, 1):
        d = load(f"ep_baseline{n}.json")
        if d:
            out += [f"torch + NCCL `all_to_all_single` baseline, {n} GPU(s): dispatch {d['dispatch_us']:.0f} µs, "
                    f"combine {d['combine_us']:.0f} µs ({d['tokens_per_s'] / 1e6:.1f} M tok/s)", ""]
    for name in sorted(glob.glob(os.path.join(P, "p2p*.json"))):
        out += p2p_table(json.load(open(name))) + [""]
    d = load("d2h.json")
    if d:
        best = max(t["mcmd_per_s"] for t in d["throughput"])
        out += [f"## GPU→CPU command queue", "",
                f"peak {best:.2f} Mcmd/s, round trip {d['latency_us']:.2f} µs (capacity {d['capacity']})", ""]
    extra = os.path.join(P, "RESULTS_extra.md")
    if os.path.exists(extra):
        out += [open(extra).read()]
    open(os.path.join(P, "RESULTS.md"), "w").write("\n".join(out) + "\n")
    print("wrote profiles/RESULTS.md,", len(out), "lines")


if __name__ == "__main__":
    main()
