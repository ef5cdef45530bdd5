#!/usr/bin/env python
"""Host link and copy-engine probes: what the box gives before any of this library's kernels run.

* H2D / D2H bandwidth of every GPU from pinned memory, optionally with the calling thread pinned to each NUMA node in
  turn (`--numa`: which node is the GPU's neighbour -- the input of `uccl_b200.net.topology.nic_for_gpu` style placement
  and of the pinned staging buffers of the P2P / proxy paths);
* device-local copy bandwidth;
* `cudaMemcpyPeerAsync` matrix (copy engines over NVLink), 1 or several streams per pair.

Device-timed with CUDA events.  Reference counterparts: experimental/misc/benchmark_pcie_bw.py,
benchmark_numa_bw.py, benchmark_memcpy_peer.py.

    python benchmarks/hostlink_bench.py [--mb 256] [--iters 20] [--numa] [--streams 4] [--out f.json]
"""
import argparse
import glob
import json
import os
import sys

import torch


def _timed(fn, iters, dev):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(dev)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.device(dev):
        s.record()
        for _ in range(iters):
            fn()
        e.record()
    torch.cuda.synchronize(dev)
    return s.elapsed_time(e) * 1e-3 / iters


def _numa_nodes():
    out = {}
    for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        try:
            cpus = set()
            for part in open(os.path.join(p, "cpulist")).read().strip().split(","):
                if "-" in part:
                    a, b = part.split("-")
                    cpus.update(range(int(a), int(b) + 1))
                elif part:
                    cpus.add(int(part))
            if cpus:
                out[int(os.path.basename(p)[4:])] = cpus
        except OSError:
            pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--numa", action="store_true", help="repeat H2D / D2H with the thread (and its first-touch pinned pages) on every NUMA node")
    ap.add_argument("--streams", type=int, default=1, help="streams per pair in the peer matrix")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if not torch.cuda.is_available():
        print(json.dumps({"unavailable": "no CUDA device"}))
        return 0
    n = torch.cuda.device_count()
    nbytes = a.mb << 20
    res = {"bytes": nbytes, "gpus": n, "host": [], "peer_GBps": None}
    nodes = _numa_nodes() if a.numa else {}
    home = os.sched_getaffinity(0)
    for g in range(n):
        dev = torch.device("cuda", g)
        d0 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        d1 = torch.empty_like(d0)
        row = {"gpu": g, "d2d_GBps": 2 * nbytes / _timed(lambda: d1.copy_(d0), a.iters, dev) / 1e9}
        for node, cpus in ([(None, None)] + sorted(nodes.items())):
            if cpus is not None:
                try:
                    os.sched_setaffinity(0, cpus & home or cpus)
                except OSError:
                    continue
            h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()  # first touch on the current node
            h.zero_()
            key = "" if node is None else f"_node{node}"
            row["h2d_GBps" + key] = nbytes / _timed(lambda: d0.copy_(h, non_blocking=True), a.iters, dev) / 1e9
            row["d2h_GBps" + key] = nbytes / _timed(lambda: h.copy_(d0, non_blocking=True), a.iters, dev) / 1e9
            del h
        os.sched_setaffinity(0, home)
        res["host"].append(row)
        print(json.dumps(row), flush=True)
    if n > 1:
        mat = [[None] * n for _ in range(n)]
        for i in range(n):
            src = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{i}")
            for j in range(n):
                if i == j:
                    continue
                dst = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{j}")
                k = max(1, a.streams)
                streams = [torch.cuda.Stream(device=i) for _ in range(k)]
                part = nbytes // k

                def go():
                    cur = torch.cuda.current_stream(i)
                    for q, st in enumerate(streams):
                        st.wait_stream(cur)
                        with torch.cuda.stream(st):
                            dst[q * part:(q + 1) * part].copy_(src[q * part:(q + 1) * part], non_blocking=True)
                    for st in streams:
                        cur.wait_stream(st)

                mat[i][j] = part * k / _timed(go, a.iters, torch.device("cuda", i)) / 1e9
                del dst
        res["peer_GBps"] = mat
        print(json.dumps({"peer_GBps": mat}), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
