"""Ad-hoc GPU diagnostics (not part of the test-suite)."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import get_world, run_ranks

def a2a_debug():
    n = 2
    comms = get_world(n)
    per = 1000
    a_ins = [torch.arange(n * per, dtype=torch.float32) + 10000 * r for r in range(n)]
    for rep in range(2):
        def prepare(c):
            x = torch.empty(n * per, device=c.device); x.copy_(a_ins[c.rank])
            o = c.empty(n * per, dtype=torch.float32); o.fill_(-1.0)
            return x, o
        outs = run_ranks(comms, prepare, lambda c, st: c.all_to_all(st[1], st[0]))
        for r, (x, o) in enumerate(outs):
            got = o.cpu()
            print(f"[a2a push rep{rep}] rank {r}: off={comms[r].native.heap_offset(o.data_ptr())} chunk0 {got[:3].tolist()} chunk1 {got[1000:1003].tolist()}")
        # search whole heap user area of each rank for the missing values
    # same with allgather for comparison
    def prep2(c):
        x = torch.empty(per, device=c.device); x.copy_(a_ins[c.rank][:per])
        o = c.empty(n * per, dtype=torch.float32); o.fill_(-1.0)
        return x, o
    outs = run_ranks(comms, prep2, lambda c, st: c.all_gather(st[1], st[0]))
    for r, (x, o) in enumerate(outs):
        got = o.cpu()
        print(f"[ag push] rank {r}: chunk0 {got[:3].tolist()} chunk1 {got[1000:1003].tolist()}")

def ep_debug():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import test_gpu_ep as T
    from uccl_b200.ep import per_token_cast_to_fp8
    n = 2
    Tn, H, K = 257, 1024, 4
    E = n * 4
    bufs = T.get_buffers(n)
    xs, idxs, ws = T.make_inputs(n, Tn, H, K, E, seed=n)
    layouts = [T.ref_layout(idxs[r], n, E) for r in range(n)]
    def fn(b):
        r = b.rank; dev = b.device
        x = xs[r].to(dev); idx = idxs[r].to(dev); w = ws[r].to(dev)
        tpr, _, tpe, in_rank, _ = b.get_dispatch_layout(idx, E)
        recv_x, recv_idx, recv_w, pe, handle, _ = b.dispatch(x, num_tokens_per_rank=tpr, is_token_in_rank=in_rank,
            num_tokens_per_expert=tpe, topk_idx=idx, topk_weights=w, use_fp8=True)
        torch.cuda.current_stream().synchronize()
        return recv_x[0].clone().cpu(), recv_x[1].clone().cpu()
    outs = T.run_threads(bufs, fn)
    for r in range(n):
        q, s = outs[r]
        rows = torch.cat([xs[src][layouts[src][2][:, r].nonzero().flatten()] for src in range(n)])
        rq, rs = per_token_cast_to_fp8(rows)
        dq = (q.view(torch.uint8) != rq.view(torch.uint8))
        ds = (s != rs)
        print(f"[ep fused] rank {r}: rows {q.shape[0]} fp8 mismatches {int(dq.sum())} ({dq.float().mean().item():.4f}), "
              f"scale mismatches {int(ds.sum())} of {ds.numel()}, max scale rel err {((s-rs).abs()/rs).max().item():.3e}")
        if dq.any():
            ij = dq.nonzero()[:8]
            for i, j in ij.tolist():
                print("   row", i, "col", j, "got", q.view(torch.uint8)[i, j].item(), "want", rq.view(torch.uint8)[i, j].item(),
                      "x", rows[i, j].item(), "scale", s[i, j // 128].item(), rs[i, j // 128].item())
            print("   mismatch rows:", dq.any(1).nonzero().flatten()[:20].tolist(), "cols/128:", sorted(set((dq.nonzero()[:, 1] // 128).tolist()))[:20])
        if ds.any():
            print("   scale mismatch rows:", ds.any(1).nonzero().flatten()[:20].tolist())

if __name__ == "__main__":
    which = sys.argv[1:] or ["a2a", "ep"]
    if "a2a" in which: a2a_debug()
    if "ep" in which: ep_debug()
