"""Prefill -> decode KV-cache hand-off BETWEEN boxes (the cross-node case of examples/kv_transfer.py).

    decode box :  python examples/kv_transfer_internode.py --server            # prints ip:port:listen_id
    prefill box:  python examples/kv_transfer_internode.py --client IP:PORT:LID

Blocks live on the GPU when there is one (staged through pinned chunks) and in host memory otherwise; the
bytes cross the network on `uccl_b200.net` (8 UDP paths, SACK/RACK recovery, Swift congestion control).
Reference role: p2p/benchmarks/benchmark_uccl.py across two hosts + the NIXL backend's xfer of KV blocks.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uccl_b200 import net  # noqa: E402
from uccl_b200.p2p import NetChannel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--server", action="store_true")
    ap.add_argument("--client", default="")
    ap.add_argument("--bind", default="")
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--block-mb", type=int, default=16)
    args = ap.parse_args()
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    eng = net.Engine(bind_ip=args.bind)
    n = (args.block_mb << 20) // 2
    if args.server:
        ch = NetChannel.listen(eng)
        ip, port, lid = ch.address
        print(f"{ip}:{port}:{lid}", flush=True)
        ch.accept()
        kv = [torch.empty(n, dtype=torch.bfloat16, device=dev) for _ in range(args.layers)]
        t0 = time.perf_counter()
        nbytes = ch.recv_tensors(kv)
        dt = time.perf_counter() - t0
        ok = all(bool((b.float() == i).all()) for i, b in enumerate(kv))
        print(f"received {nbytes / 1e6:.1f} MB in {dt * 1e3:.1f} ms ({nbytes * 8e-9 / dt:.2f} Gb/s), payload {'ok' if ok else 'CORRUPT'}")
        ch.close()
        return 0 if ok else 1
    ip, port, lid = args.client.split(":")
    ch = NetChannel.connect(eng, (ip, int(port), int(lid)))
    kv = [torch.full((n,), float(i), dtype=torch.bfloat16, device=dev) for i in range(args.layers)]
    t0 = time.perf_counter()
    nbytes = ch.send_tensors(kv)
    dt = time.perf_counter() - t0
    print(f"sent {nbytes / 1e6:.1f} MB in {dt * 1e3:.1f} ms ({nbytes * 8e-9 / dt:.2f} Gb/s); flow {eng.flow_stats(ch.flow)}")
    ch.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
