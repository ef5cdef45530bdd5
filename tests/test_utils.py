"""CPU unit tests of the host utilities (rings, pool, histogram, seqnos, congestion control) --
the plain-`main` unit tests of the reference (include/util/util_test.cc, collective/*/timely_test.cc)."""
from uccl_b200 import _native

U = _native.C().util


def test_spsc_ring_fifo_and_capacity():
    r = U.SpscRing(8)
    assert r.capacity == 8
    for i in range(8):
        assert r.push(i)
    assert not r.push(99)
    assert [r.pop() for _ in range(8)] == list(range(8))
    assert r.pop() is None
    for rnd in range(100):  # wrap-around
        assert r.push(rnd) and r.pop() == rnd


def test_mpmc_ring_concurrent_stress():
    r = U.MpmcRing(256)
    producers, consumers, per = 4, 3, 20000
    s, c = r.stress(producers, consumers, per)
    assert c == producers * per
    assert s == producers * per * (per + 1) // 2
    assert r.pop() is None


def test_shared_pool_recycles():
    p = U.SharedPool(64)
    for i in range(40):
        assert p.release_global(i)
    got = [p.get() for _ in range(40)]
    assert sorted(got) == list(range(40))
    assert p.get() is None
    for v in got:
        p.put(v)
    assert sorted(p.get() for _ in range(40)) == list(range(40))


def test_latency_hist_percentiles():
    h = U.LatencyHist()
    for v in range(1, 10001):
        h.record(v)
    assert h.count() == 10000 and h.min() == 1 and h.max() == 10000
    assert abs(h.mean() - 5000.5) < 1e-6
    for p, exp in ((50, 5000), (90, 9000), (99, 9900)):
        got = h.percentile(p)
        assert exp <= got <= exp * 1.07, (p, got)  # <= 1/16 relative bucket error
    assert "p99" in h.summary()


def test_seqno_wraparound():
    assert U.seqno_less(16, 65530, 5)      # 65530 is "before" 5 after wrap
    assert not U.seqno_less(16, 5, 65530)
    assert U.seqno_less(8, 250, 3)
    assert U.seqno_less(32, 1, 2)


def test_timely_reacts_to_rtt_gradient():
    t = U.Timely()
    r0 = t.rate_gbps()
    rtt = 10.0
    for _ in range(30):  # rising RTT inside [t_low, t_high] -> multiplicative decrease
        rtt += 1.5
        t.on_rtt(rtt)
    assert t.rate_gbps() < 0.5 * r0
    low = t.rate_gbps()
    for _ in range(200):  # RTT back at the floor -> additive (then hyper-active) increase
        t.on_rtt(3.0)
    assert t.rate_gbps() > low
    assert t.pacing_delay_us(1 << 20) > 0


def test_swift_window_tracks_target_delay():
    s = U.Swift()
    c0 = s.cwnd()
    now = 0.0
    for _ in range(50):
        now += 10.0
        s.on_ack(2.0, 1.0, now, 10.0)  # below target -> grow
    assert s.cwnd() > c0
    grown = s.cwnd()
    for _ in range(50):
        now += 10.0
        s.on_ack(500.0, 1.0, now, 10.0)  # far above target -> shrink (at most once per RTT)
    assert s.cwnd() < grown * 0.1
    assert s.target_delay_us() > 8.0  # flow scaling raises the target for tiny windows


def test_eqds_pacer_is_fair_and_rate_limited():
    p = U.EqdsPacer()
    for sender in range(4):
        p.add_demand(sender, 64 << 20)
    t, total = 0.0, 0
    for _ in range(2000):
        t += 1.0  # 1 us ticks; 900 GB/s -> 900 KB per tick ~ 13 credits of 64 KiB
        for sender, nbytes in p.tick(t):
            total += nbytes
            p.on_data(sender, nbytes)
    grants = [p.granted(s) for s in range(4)]
    assert sum(grants) == total
    assert max(grants) - min(grants) <= 2 * 65536   # round-robin fairness
    assert total <= 900e9 * 2000e-6 * 1.02          # never above line rate
    assert total >= min(4 * (64 << 20), 900e9 * 2000e-6 * 0.9)
