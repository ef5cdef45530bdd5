"""GPU -> CPU command queue + proxy: device-initiated copy-engine writes with an ordered remote
counter update, notifications, and the queue microbenchmarks (the reference exercises its FIFO the
same way: ep/bench/fifo, ep/tests/{gpu_to_cpu,batched_gpu_to_cpu}_bench.cu)."""
import time

import pytest
import torch

from helpers import get_world
from uccl_b200.ep import Proxy

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_device_initiated_write_with_signal():
    comms = get_world(2)
    c0, c1 = comms
    with torch.cuda.device(c0.device):
        src = c0.empty(1 << 16, dtype=torch.float32)
        src.copy_(torch.arange(1 << 16, dtype=torch.float32))
    with torch.cuda.device(c1.device):
        dst = c1.empty(1 << 16, dtype=torch.float32)
        dst.zero_()
        sig = c1.empty(2, dtype=torch.int64)
        sig.zero_()
    for c in comms:
        torch.cuda.synchronize(c.device)
    p = Proxy(c0, capacity=256)
    try:
        dst_off = c1._c.heap_offset(dst.data_ptr())
        sig_off = c1._c.heap_offset(sig.data_ptr())
        with torch.cuda.device(c0.device):
            for i in range(3):  # three writes, each followed by its signal
                p.device_write(1, src, dst_off, signal_offset=sig_off, signal_value=5)
            p.device_notify(tag=7, value=1234)
            torch.cuda.current_stream().synchronize()
        p.drain()
        torch.cuda.synchronize(c1.device)
        assert torch.equal(dst.cpu(), torch.arange(1 << 16, dtype=torch.float32))
        assert sig[0].item() == 15
        notes = []
        t0 = time.time()
        while not notes and time.time() - t0 < 5:
            notes = p.poll_notifications()
        assert notes == [(7, 1234)]
        st = p.stats()
        assert st["writes"] == 3 and st["atomics"] == 3 and st["notifies"] == 1 and st["bytes"] == 3 * 4 * (1 << 16)
    finally:
        p.stop()


def test_queue_microbench_and_flow_control():
    c = get_world(2)[0]
    p = Proxy(c, capacity=64)  # far fewer slots than commands: producers must block on the consumer
    try:
        with torch.cuda.device(c.device):
            rate = p.bench_throughput(blocks=4, threads=64, per_thread=32)
            lat = p.bench_latency(iters=200)
        st = p.stats()
        assert st["nops"] == 4 * 64 * 32 + 200
        assert rate > 1e4 and 0.5 < lat < 5000, (rate, lat)
        print(f"d2h queue: {rate / 1e6:.2f} Mcmd/s, round trip {lat:.1f} us")
    finally:
        p.stop()
