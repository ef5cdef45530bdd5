"""Latency / bandwidth of the inter-node datagram transport (``uccl_b200.net``).

Single process (two engines over loopback) by default; with ``--server`` / ``--client IP:PORT:LID`` the two
ends run on different hosts -- the shape of the reference's ``p2p/benchmarks`` and ``collective/rdma``
transport tests.  Prints one row per message size and a JSON summary.

    python benchmarks/net_bench.py --paths 8 --payload 8192 --cc swift
    python benchmarks/net_bench.py --server                      # prints ip:port:listen_id
    python benchmarks/net_bench.py --client 10.0.0.2:40123:1
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uccl_b200 import net  # noqa: E402


def sizes(lo, hi):
    s = lo
    while s <= hi:
        yield s
        s *= 4


def run_pair(send_e, send_f, recv_e, recv_f, args):
    rows = []
    for nbytes in sizes(args.min_bytes, args.max_bytes):
        x = torch.empty(nbytes, dtype=torch.uint8).random_(0, 255)
        y = torch.empty_like(x)
        iters = max(3, min(args.iters, (512 << 20) // nbytes))
        for _ in range(3):
            w = recv_e.irecv(recv_f, y)
            send_e.send(send_f, x)
            w.wait()
        t0 = time.perf_counter()
        for _ in range(iters):
            w = recv_e.irecv(recv_f, y)
            send_e.send(send_f, x)
            w.wait()
        dt = (time.perf_counter() - t0) / iters
        assert torch.equal(x, y)
        rows.append({"bytes": nbytes, "us": dt * 1e6, "gbps": nbytes * 8e-9 / dt})
        print(f"{nbytes:>12} B {dt * 1e6:>12.1f} us {nbytes * 8e-9 / dt:>9.2f} Gb/s", flush=True)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", default="127.0.0.1")
    ap.add_argument("--paths", type=int, default=8)
    ap.add_argument("--payload", type=int, default=8192)
    ap.add_argument("--cc", default="swift", choices=list(net.CC))
    ap.add_argument("--drop", type=float, default=0.0, help="injected loss probability per datagram")
    ap.add_argument("--min-bytes", type=int, default=4096)
    ap.add_argument("--max-bytes", type=int, default=64 << 20)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--busy-poll", action="store_true")
    ap.add_argument("--server", action="store_true")
    ap.add_argument("--client", default="")
    args = ap.parse_args()
    mk = lambda: net.Engine(bind_ip=args.bind, paths=args.paths, payload=args.payload, cc=args.cc,  # noqa: E731
                            drop_prob=args.drop, busy_poll=args.busy_poll)
    if args.server:
        e = mk()
        lid = e.listen()
        print(f"{e.address}:{e.port}:{lid}", flush=True)
        f = e.accept(lid, timeout_ms=600000)
        buf = torch.empty(args.max_bytes, dtype=torch.uint8)
        n = torch.zeros(1, dtype=torch.int64)
        while True:  # echo protocol: 8-byte size, then that many payload messages until size 0
            e.recv(f, n)
            if int(n) == 0:
                break
            e.recv(f, buf[: int(n)])
            e.send(f, n)
        return
    if args.client:
        ip, port, lid = args.client.split(":")
        e = mk()
        f = e.connect(ip, int(port), int(lid))
        rows = []
        ack = torch.zeros(1, dtype=torch.int64)
        for nbytes in sizes(args.min_bytes, args.max_bytes):
            x = torch.empty(nbytes, dtype=torch.uint8)
            iters = max(3, min(args.iters, (512 << 20) // nbytes))
            t0 = time.perf_counter()
            for _ in range(iters):
                e.send(f, torch.tensor([nbytes]))
                e.send(f, x)
                e.recv(f, ack)
            dt = (time.perf_counter() - t0) / iters
            rows.append({"bytes": nbytes, "us": dt * 1e6, "gbps": nbytes * 8e-9 / dt})
            print(f"{nbytes:>12} B {dt * 1e6:>12.1f} us {nbytes * 8e-9 / dt:>9.2f} Gb/s", flush=True)
        e.send(f, torch.tensor([0]))
        print(json.dumps({"bench": "net_client", "rows": rows, "flow": e.flow_stats(f)}))
        return
    a, b = mk(), mk()
    lid = b.listen()
    box = {}
    t = threading.Thread(target=lambda: box.setdefault("f", b.accept(lid)))
    t.start()
    fa = a.connect(args.bind, b.port, lid)
    t.join()
    rows = run_pair(a, fa, b, box["f"], args)
    st = a.flow_stats(fa)
    print(json.dumps({"bench": "net_loopback", "paths": args.paths, "payload": args.payload, "cc": args.cc,
                      "drop": args.drop, "rows": rows,
                      "flow": {k: st[k] for k in ("tx_pkts", "fast_rexmit", "rto_rexmit", "srtt_us", "cwnd", "path_tx")}}))


if __name__ == "__main__":
    main()
