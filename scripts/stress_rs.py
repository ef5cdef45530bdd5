"""Stress the staged (plain-buffer) reduce_scatter variants with tiny stages (many chunks) in a
single-process world of 8 ranks; prints the first mismatch in detail.  Debug aid."""
import sys
import os

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from helpers import get_world, run_ranks


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    comms = get_world(n, heap_mb=160, stage_mb=1, max_ctas=4)
    for c in comms:
        c.set_xchg_ll_max(-1)
    bad_total = 0
    for it in range(iters):
        for dtype, op, count in ((torch.float32, "sum", (1 << 18) + 4), (torch.bfloat16, "max", 70001)):
            g = torch.Generator().manual_seed(it * 7 + count)
            ins = [(torch.randn(n * count, generator=g) * 2).to(dtype) for _ in range(n)]
            ref = torch.stack([x.double() for x in ins])
            exp = (ref.sum(0) if op == "sum" else ref.max(0).values).view(n, count)
            for push in (False, True):
                for c in comms:
                    c.set_rs_push(push)
                outs = run_ranks(comms, lambda c: (ins[c.rank].to(c.device), torch.zeros(count, dtype=dtype, device=c.device)),
                                 lambda c, st: c.reduce_scatter(st[1], st[0], op))
                for r, (_, o) in enumerate(outs):
                    got = o.cpu().double()
                    tol = dict(rtol=2e-2, atol=6e-2) if dtype == torch.bfloat16 else dict(rtol=1e-5, atol=2e-5)
                    bad = (~torch.isclose(got, exp[r], **tol)).nonzero().flatten()
                    if bad.numel():
                        bad_total += 1
                        es = 2 if dtype == torch.bfloat16 else 4
                        print(f"it {it} push={push} {dtype} rank {r}: {bad.numel()} bad elems, byte range "
                              f"[{bad[0].item() * es}, {bad[-1].item() * es}], first idx {bad[:6].tolist()}, got "
                              f"{got[bad[:3]].tolist()} want {exp[r][bad[:3]].tolist()}", flush=True)
                        # which source is missing? compare against partial sums
                        if op == "sum":
                            i = bad[0].item()
                            contrib = [ins[s][r * count + i].item() for s in range(n)]
                            print("   contributions", contrib, "sum", sum(contrib), flush=True)
        if bad_total > 6:
            break
    print("stress done, mismatching (iter,rank) cases:", bad_total)


if __name__ == "__main__":
    main()
