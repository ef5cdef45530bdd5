"""``python -m uccl_b200 [--json]``: what is installed and what the machine looks like -- the first thing to attach to
a bug report (the reference spreads this over scripts/, `nvidia-smi topo` calls in its READMEs and UCCL_DEBUG=INFO logs).
Needs no GPU; with GPUs it adds the device list, peer-access matrix and the NVLS (multicast) capability."""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys


def collect() -> dict:
    import torch

    import uccl_b200
    from uccl_b200 import _build, _native

    info = {"version": uccl_b200.__version__, "python": sys.version.split()[0], "torch": torch.__version__,
            "torch_cuda": torch.version.cuda, "arch_flags": " ".join(_build.ARCH_FLAGS), "nvcc": _build.NVCC,
            "module": str(_build.module_path()), "module_built": os.path.exists(_build.module_path()),
            "nccl_shim": uccl_b200.nccl_shim_path(), "nccl_shim_built": os.path.exists(uccl_b200.nccl_shim_path()),
            "nccl_net_plugin": uccl_b200.nccl_plugin_path(),
            "nccl_net_plugin_built": os.path.exists(uccl_b200.nccl_plugin_path())}
    try:
        C = _native.C()
        info["native"] = {"max_ranks": int(C.MAX_RANKS), "ll_max_bytes": int(C.LL_MAX_BYTES)}
    except Exception as e:  # noqa: BLE001
        info["native"] = {"error": f"{type(e).__name__}: {e}"}
    info["env"] = {k: v for k, v in sorted(os.environ.items()) if k.startswith(("UCCL_", "NCCL_", "CUDA_VISIBLE"))}
    gpus = []
    if torch.cuda.is_available():
        n = torch.cuda.device_count()
        for i in range(n):
            p = torch.cuda.get_device_properties(i)
            gpus.append({"index": i, "name": p.name, "cc": f"{p.major}.{p.minor}", "sms": p.multi_processor_count,
                         "memory_GiB": round(p.total_memory / 2 ** 30, 1)})
        info["peer_access"] = [[bool(i == j or torch.cuda.can_device_access_peer(i, j)) for j in range(n)]
                               for i in range(n)]
        try:
            from cuda import cuda as drv  # cuda-python, optional

            drv.cuInit(0)
            err, v = drv.cuDeviceGetAttribute(drv.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, 0)
            info["multicast_supported"] = bool(v) if int(err) == 0 else None
        except Exception:  # noqa: BLE001
            info["multicast_supported"] = None
        smi = shutil.which("nvidia-smi")
        if smi:
            r = subprocess.run([smi, "topo", "-m"], capture_output=True, text=True, timeout=30)
            info["topology"] = r.stdout.strip().splitlines()[: n + 2]
    info["gpus"] = gpus
    try:
        from uccl_b200.utils import SmPartition

        ok, why = SmPartition.supported(0 if gpus else None)
        info["sm_partitions"] = "supported" if ok else f"unavailable ({why})"
    except Exception as e:  # noqa: BLE001
        info["sm_partitions"] = f"unavailable ({type(e).__name__}: {e})"
    try:
        from uccl_b200.net import topology

        info["nics"] = [str(x) for x in topology.list_nics()] if hasattr(topology, "list_nics") else None
    except Exception:  # noqa: BLE001
        info["nics"] = None
    return info


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    info = collect()
    if "--json" in argv:
        print(json.dumps(info, indent=1))
        return 0
    print(f"uccl_b200 {info['version']}  (python {info['python']}, torch {info['torch']}, CUDA {info['torch_cuda']})")
    print(f"  build: {info['arch_flags']}  nvcc={info['nvcc']}")
    for k in ("module", "nccl_shim", "nccl_net_plugin"):
        state = "built" if info[k + "_built"] else 'MISSING: python -c "import uccl_b200; uccl_b200.build()"'
        print(f"  {k:16s} {info[k]}  [{state}]")
    print(f"  native: {info['native']}")
    if info["gpus"]:
        for g in info["gpus"]:
            print(f"  GPU {g['index']}: {g['name']} cc {g['cc']}, {g['sms']} SMs, {g['memory_GiB']} GiB")
        full = all(all(r) for r in info["peer_access"])
        print(f"  peer access: {'full mesh' if full else info['peer_access']}; NVLS multicast: {info['multicast_supported']}")
        print(f"  SM partitions (green contexts): {info['sm_partitions']}")
        for line in info.get("topology", []):
            print("   ", line)
    else:
        print("  no CUDA device: host backends only (host communicators, ep host Buffer, p2p.Endpoint(-1), ukernel host worker)")
    if info["env"]:
        print("  environment:", " ".join(f"{k}={v}" for k, v in info["env"].items()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
