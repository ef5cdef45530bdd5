#!/usr/bin/env python
"""profiles/RESULTS_r2.md from the round-2 JSON files in profiles/ (no GPU needed).

    python scripts/make_results_r2.py
"""
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
    except json.JSONDecodeError:
        for line in reversed(txt.splitlines()):
            if line.startswith("{"):
                return json.loads(line)
    return None


def ep_sweep(out, n):
    d = load(f"ep_sweep_{n}xB200.json")
    if not d:
        return
    rows = [r for r in d["rows"] if not r.get("ll")]
    sms = sorted({r["sms"] for r in rows})
    out += [f"### EP dispatch / combine sweep, {n}×B200 ({d['tokens']} tokens/rank, hidden {d['hidden']}, top-{d['topk']}, "
            f"{d['experts']} experts; cached handle, CUDA events, 256 MiB L2 flush per call, max over ranks)", "",
            "| CTAs | kernels | dispatch bf16→fp8 fused | dispatch bf16 | combine bf16 |", "|---:|---|---:|---:|---:|"]
    for s_ in sms:
        for impl in ("reg", "tma"):
            f8 = next((r for r in rows if r["sms"] == s_ and r.get("impl", "reg") == impl and r["mode"] == "fp8_fused"), None)
            bf = next((r for r in rows if r["sms"] == s_ and r.get("impl", "reg") == impl and r["mode"] == "bf16"), None)
            if not f8 and not bf:
                continue
            c = lambda r, k, g: f"{r[k]:.0f} µs ({r[g]:.0f} GB/s)" if r and k in r else "-"  # noqa: E731
            out.append(f"| {s_} | {'register' if impl == 'reg' else 'TMA'} | {c(f8, 'dispatch_us', 'dispatch_GBps')} | "
                       f"{c(bf, 'dispatch_us', 'dispatch_GBps')} | {c(bf, 'combine_us', 'combine_GBps')} |")
    ll = [r for r in d["rows"] if r.get("ll")]
    for r in ll:
        extra = f", back-to-back dispatch+combine pair {r['pair_back_to_back_us']:.1f} µs" if "pair_back_to_back_us" in r else ""
        out.append("")
        out.append(f"Low-latency, {r['tokens']} tokens/rank, fp8={r['use_fp8']}: dispatch {r['dispatch_us']:.1f} µs, "
                   f"combine {r['combine_us']:.1f} µs (each timed alone after an L2 flush){extra}")
    out.append("")


def anchor(out, n):
    d = load(f"deepep_anchor_{n}xB200.json")
    b = load(f"bench{n}.json")
    if not d:
        return
    a = d.get("diagnostic", {}).get("vendored_upstream_deepep_intranode", {})
    if "best_us" not in a:
        return
    out += [f"### Same-box anchor, {n}×B200: upstream DeepEP (reference/thirdparty/DeepEP, unmodified) at num_sms = 24", "",
            "| | upstream DeepEP, best NVL chunk | this library, 24 CTAs | ratio |", "|---|---:|---:|---:|"]
    ours = next((r for r in (b or {}).get("sm_sweep", []) if r["num_sms"] == 24), None)
    if ours:
        out.append(f"| dispatch to fp8 | {a['best_us']['dispatch_fp8']:.0f} µs (+ {a['fp8_cast_torch_us']:.0f} µs torch cast before it) | "
                   f"{ours['dispatch_us']:.0f} µs (cast fused) | {a['best_us']['dispatch_fp8'] / ours['dispatch_us']:.2f}× "
                   f"({a['dispatch_fp8_incl_cast_us'] / ours['dispatch_us']:.2f}× incl. cast) |")
        out.append(f"| combine bf16 | {a['best_us']['combine_bf16']:.0f} µs | {ours['combine_us']:.0f} µs | "
                   f"{a['best_us']['combine_bf16'] / ours['combine_us']:.2f}× |")
        step = ours["ms_per_step"] * 1e3
        out.append(f"| dispatch + combine step | {a['step_us_fp8_dispatch_plus_bf16_combine']:.0f} µs ({a['step_us_incl_cast']:.0f} µs incl. cast) | "
                   f"{step:.0f} µs | {a['step_us_fp8_dispatch_plus_bf16_combine'] / step:.2f}× ({a['step_us_incl_cast'] / step:.2f}×) |")
    out.append(f"| layout | {a['layout_us']:.1f} µs | (multi-CTA layout kernel: see ncu launch list) | |")
    out.append("")
    out.append("DeepEP by NVL chunk size (µs): " + "; ".join(
        f"{k}: " + ", ".join(f"{c}={v:.0f}" for c, v in vals.items()) for k, vals in a["by_nvl_chunk_us"].items()))
    out.append("")


def bench(out):
    out += ["### bench.py (driver contract; headline = the reference's 24-SM budget for N > 1)", "",
            "| GPUs | headline CTAs | tokens/s | step | dispatch / combine | best of SM sweep | e2e tokens/s |", "|---:|---:|---:|---:|---:|---:|---:|"]
    for n in (1, 2, 4, 8):
        b = load(f"bench{n}.json")
        if not b or "sm_sweep" not in b:
            continue
        best = b["best"]
        out.append(f"| {n} | {b['config']['num_sms']} ({b['config']['kernels']['dispatch']}) | {b['value'] / 1e6:.1f} M | {b['ms_per_step'] * 1e3:.0f} µs | "
                   f"{b['dispatch_us']:.0f} / {b['combine_us']:.0f} µs | {best['tokens_per_s'] / 1e6:.1f} M @ {best['num_sms']} CTAs | "
                   f"{b['e2e']['value'] / 1e6:.1f} M |")
    out.append("")


def ar_plain(out, n):
    d = load(f"ar_plain_{n}xB200.json")
    if not d:
        return
    out += [f"### AllReduce on ordinary (cudaMalloc) tensors, bf16 sum, {n}×B200 -- µs (bus GB/s)", "",
            "| bytes | auto (picked) | staged_nvls | staged_pipe | NCCL |", "|---:|---:|---:|---:|---:|"]
    for r in d["rows"]:
        c = lambda k: f"{r[k]['us']:.0f} ({r[k]['busbw_GBps']:.0f})" if k in r and "us" in r[k] else "-"  # noqa: E731
        out.append(f"| {r['bytes']} | {c('auto')} [{r.get('auto', {}).get('picked', '')}] | {c('staged_nvls')} | {c('staged_pipe')} | {c('nccl')} |")
    out.append("")


def p2p(out):
    d = load("p2p_2xB200.json")
    if not d:
        return
    out += ["### P2P engine GPU0→GPU1 (2×B200), host-timed issue + completion like the reference's benchmark_uccl.py", "",
            "| bytes | blocks | write | prepared write | read | async ×4 | dual / dir | one cudaMemcpyPeer per block | write vs memcpy |",
            "|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for r in d["rows"]:
        g = lambda k: f"{r[k]['GBps']:.0f} GB/s ({r[k]['us']:.0f} µs)" if k in r else "-"  # noqa: E731
        out.append(f"| {r['bytes']} | {r['blocks']} | {g('write')} | {g('prepared_write')} | {g('read')} | {r['async']['GBps']:.0f} GB/s | "
                   f"{r['dual']['GBps_per_direction']:.0f} GB/s | {g('memcpy_per_block')} | {r['speedup_write_vs_memcpy']:.2f}× "
                   f"({r.get('speedup_prepared_vs_memcpy', 0):.2f}× prepared) |")
    out.append("")


def ddp(out):
    for n in (2, 4, 8):
        path = os.path.join(P, f"ddp_resnet50_{n}xB200.jsonl")
        if not os.path.exists(path):
            continue
        rows = [json.loads(l) for l in open(path) if l.strip()]
        out += [f"### DDP ResNet-50 (bf16 autocast, batch {rows[0]['batch_per_gpu']}/GPU, synthetic ImageNet-shaped data), {n}×B200", "",
                "| backend | img/s | ms/step |", "|---|---:|---:|"]
        for r in rows:
            out.append(f"| {r['backend']} | {r['img_per_s']:.0f} | {r['ms_per_step']:.1f} |")
        out.append("")


def nccl_tests(out):
    for n in (2, 4, 8):
        path = os.path.join(P, f"nccl_tests_{n}xB200.md")
        if os.path.exists(path):
            out += [f"### nccl-tests (reference/thirdparty/nccl-tests, unmodified) on {n}×B200: system NCCL 2.27.3 vs the drop-in preloaded, `-c 1`",
                    "", f"Full tables: `profiles/nccl_tests_{n}xB200.md`.", ""]


def main():
    out = ["# RESULTS — round 2 (B200, sm_100a)", "",
           "Device-timed (CUDA events) and max over ranks unless a table says otherwise; generated by `scripts/make_results_r2.py` "
           "from the JSON files in this directory.  Round-1 tables (collectives on symmetric buffers etc.): `RESULTS.md`.", ""]
    bench(out)
    for n in (8, 4, 2):
        anchor(out, n)
    for n in (8, 4, 2, 1):
        ep_sweep(out, n)
    for n in (8, 4):
        ar_plain(out, n)
    nccl_tests(out)
    ddp(out)
    p2p(out)
    with open(os.path.join(P, "RESULTS_r2.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote profiles/RESULTS_r2.md", len(out), "lines")


if __name__ == "__main__":
    main()
