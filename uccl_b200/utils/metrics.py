"""Prometheus exporter for the runtime counters (kernel launches of a communicator, P2P engine bytes /
transfers, ukernel worker tasks, CPU proxy commands, symmetric-heap headroom).

The reference only prints periodic status lines and ships shell monitors; a serving
deployment wants the same numbers scrapeable::

    from uccl_b200.utils.metrics import MetricsExporter
    exp = MetricsExporter(rank=comm.rank)
    exp.watch_communicator(comm); exp.watch_endpoint(ep); exp.watch_ukernel(uk_comm); exp.watch_net_engine(engine)
    exp.start_http_server(9400 + comm.rank)      # or: text = exp.render()
"""
from __future__ import annotations

from typing import Callable, Dict, List, Tuple


class MetricsExporter:
    def __init__(self, rank: int = 0, namespace: str = "uccl_b200"):
        from prometheus_client import CollectorRegistry

        self.rank = str(rank)
        self.ns = namespace
        self.registry = CollectorRegistry()
        self._sources: List[Tuple[str, Callable[[], Dict[str, float]]]] = []
        self.registry.register(self)

    # ---- prometheus collector protocol
    def collect(self):
        from prometheus_client.core import GaugeMetricFamily

        for prefix, fn in self._sources:
            try:
                vals = fn()
            except Exception:  # a stopped component must not break the scrape
                continue
            for k, v in vals.items():
                g = GaugeMetricFamily(f"{self.ns}_{prefix}_{k}", f"{prefix} {k}", labels=["rank"])
                g.add_metric([self.rank], float(v))
                yield g

    # ---- sources
    def watch_communicator(self, comm, name: str = "comm"):
        self._sources.append((name, lambda: {"kernel_launches": comm.native.launches,
                                              "heap_free_bytes": comm.native.heap_free_bytes,
                                              "error_word": comm.native.error_word}))

    def watch_endpoint(self, ep, name: str = "p2p"):
        self._sources.append((name, lambda: {k: v for k, v in ep.stats().items()}))

    def watch_ukernel(self, uk, name: str = "ukernel"):
        self._sources.append((name, lambda: {k: v for k, v in uk.stats().items()}))

    def watch_proxy(self, proxy, name: str = "proxy"):
        self._sources.append((name, lambda: {k: v for k, v in proxy.stats().items()}))

    def watch_net_engine(self, engine, name: str = "net"):
        """Inter-node transport: datagrams, retransmissions, drops, engine-loop activity (``uccl_b200.net.Engine``)."""
        self._sources.append((name, lambda: {k: v for k, v in engine.stats().items()}))

    def watch_net_communicator(self, nc, name: str = "net_flow"):
        """Per-peer flow health of a ``NetCommunicator``: srtt, cwnd, retransmissions, quarantined paths."""

        def fn():
            out = {}
            for peer, f in nc.flows.items():
                st = nc.engine.flow_stats(f) or {}
                for k in ("srtt_us", "cwnd", "fast_rexmit", "rto_rexmit", "path_bans", "tx_bytes", "rx_bytes"):
                    out[f"peer{peer}_{k}"] = st.get(k, 0)
            return out

        self._sources.append((name, fn))

    def watch(self, name: str, fn: Callable[[], Dict[str, float]]):
        self._sources.append((name, fn))

    # ---- output
    def render(self) -> str:
        from prometheus_client import generate_latest

        return generate_latest(self.registry).decode()

    def start_http_server(self, port: int, addr: str = "127.0.0.1"):
        from prometheus_client import start_http_server

        return start_http_server(port, addr=addr, registry=self.registry)
