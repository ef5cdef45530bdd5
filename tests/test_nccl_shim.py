"""Builds and runs the C++ NCCL-API test against libuccl_b200_nccl.so on the host backend."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_test(tmp_path):
    from uccl_b200 import _build

    _build.build()
    shim = _build.nccl_shim_path()
    assert shim.exists()
    exe = tmp_path / "nccl_api_test"
    cmd = ["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests/cpp/nccl_api_test.cc"), "-I/usr/include",
           "-I/usr/local/cuda/include", "-L" + str(shim.parent), "-luccl_b200_nccl", "-Wl,-rpath," + str(shim.parent),
           "-o", str(exe)]
    subprocess.run(cmd, check=True)
    return exe


@pytest.mark.timeout(300)
def test_nccl_api_world2_cpu(tmp_path):
    exe = _build_test(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=240)
    sys.stdout.write(r.stdout)
    sys.stderr.write(r.stderr)
    assert r.returncode == 0, r.stderr
    assert "nccl_api_test: OK" in r.stdout


@pytest.mark.timeout(300)
def test_nccl_api_spans_boxes_cpu(tmp_path):
    """ncclCommInitRank with 4 ranks and a box size of 2: the drop-in builds a MultiComm (native communicator per
    box + datagram rails) and every NCCL collective, PreMulSum and grouped send/recv work across the boxes."""
    from uccl_b200 import _build

    _build.build()
    shim = _build.nccl_shim_path()
    exe = tmp_path / "nccl_multibox_test"
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests/cpp/nccl_multibox_test.cc"), "-I/usr/include",
                    "-I/usr/local/cuda/include", "-L" + str(shim.parent), "-luccl_b200_nccl",
                    "-Wl,-rpath," + str(shim.parent), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=240)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "nccl_multibox_test: OK" in r.stdout


def test_exported_symbols():
    from uccl_b200 import _build

    _build.build()
    out = subprocess.run(["nm", "-D", str(_build.nccl_shim_path())], capture_output=True, text=True).stdout
    for sym in ("ncclAllReduce", "ncclAllGather", "ncclReduceScatter", "ncclBroadcast", "ncclReduce", "ncclSend",
                "ncclRecv", "ncclAllToAll", "ncclGroupStart", "ncclGroupEnd", "ncclCommInitRank", "ncclCommInitAll",
                "ncclCommSplit", "ncclGetUniqueId", "ncclMemAlloc", "ncclMemFree", "ncclCommRegister"):
        assert f" T {sym}\n" in out, sym


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("n", [2, 4])
def test_nccl_api_gpu(tmp_path, n):
    """NCCL C API on the GPU: InitAll, staged + zero-copy allreduce, allgather, reduce-scatter,
    broadcast, native grouped send/recv."""
    from uccl_b200 import _build

    _build.build()
    shim = _build.nccl_shim_path()
    exe = tmp_path / "nccl_gpu_test"
    cmd = ["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests/cpp/nccl_gpu_test.cc"), "-I/usr/include",
           "-I/usr/local/cuda/include", "-L" + str(shim.parent), "-luccl_b200_nccl", "-Wl,-rpath," + str(shim.parent),
           "-L/usr/local/cuda/lib64", "-lcudart", "-lpthread", "-o", str(exe)]
    subprocess.run(cmd, check=True)
    import torch

    env = dict(os.environ)
    if torch.cuda.device_count() < n:  # virtual ranks share one GPU: give every stream its own hardware queue
        env["CUDA_DEVICE_MAX_CONNECTIONS"] = "32"
    r = subprocess.run([str(exe), str(n)], capture_output=True, text=True, timeout=240, env=env)
    sys.stdout.write(r.stdout)
    sys.stderr.write(r.stderr)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "nccl_gpu_test: OK" in r.stdout


@pytest.mark.timeout(300)
def test_nccl_fallback_forwarding_cpu(tmp_path):
    """UCCL_B200_NCCL_FALLBACK_{LIB,OPS,MIN_BYTES}: listed operations go to a dlopen'd libnccl (here the stand-in
    tests/cpp/fake_nccl.cc), everything else stays native; rank 0's id reaches every rank; a misspelt operation
    fails communicator creation (reference: lite's dlopen fallback + force list, nccl.cu:707-860,1866-1872)."""
    from uccl_b200 import _build

    _build.build()
    shim = _build.nccl_shim_path()
    fake = tmp_path / "libfake_nccl.so"
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", os.path.join(ROOT, "tests/cpp/fake_nccl.cc"),
                    "-I/usr/include", "-I/usr/local/cuda/include", "-o", str(fake)], check=True)
    exe = tmp_path / "nccl_fallback_test"
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests/cpp/nccl_fallback_test.cc"), "-I/usr/include",
                    "-I/usr/local/cuda/include", "-L" + str(shim.parent), "-luccl_b200_nccl",
                    "-Wl,-rpath," + str(shim.parent), "-ldl", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe), str(fake)], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "nccl_fallback_test: OK" in r.stdout, r.stdout + r.stderr[-3000:]


def test_drop_in_covers_what_torch_and_the_newest_header_need():
    """Symbol audit (role of the reference's lite/nccl/audit_nccl.cc): every nccl* symbol torch's CUDA library imports,
    and every entry point of the newest nccl.h found on this machine, must be exported by the drop-in -- a preloaded
    drop-in that lacks one would let that call fall through to another libnccl with a foreign communicator handle."""
    import glob
    import re

    import torch

    from uccl_b200 import _build

    _build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", str(_build.nccl_shim_path())], capture_output=True, text=True).stdout
    ours = set(re.findall(r"\bp?nccl[A-Za-z0-9_]+", out))
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib", "libtorch_cuda.so")
    if os.path.exists(tlib):
        und = subprocess.run(["nm", "-D", "--undefined-only", tlib], capture_output=True, text=True).stdout
        need = set(re.findall(r"\bnccl[A-Za-z0-9_]+", und))
        assert need, "torch imports no nccl symbols?"
        assert not (need - ours), f"torch needs {sorted(need - ours)}"
    headers = glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "nccl", "include", "nccl.h"))
    headers.append("/usr/include/nccl.h")
    for h in headers:
        if not os.path.exists(h):
            continue
        api = set(re.findall(r"^(?:ncclResult_t|const char\*)\s+(nccl[A-Za-z0-9_]+)\s*\(", open(h).read(), flags=re.M))
        assert len(api) > 30
        assert not (api - ours), f"{h}: not exported: {sorted(api - ours)}"


def test_drop_in_covers_vllm_pynccl_bindings():
    """vLLM loads libnccl with ctypes (VLLM_NCCL_SO_PATH selects the file): every function its wrapper binds must be
    an export of the drop-in.  Skipped when vLLM is not installed."""
    import importlib.util
    import re

    from uccl_b200 import _build

    spec = importlib.util.find_spec("vllm")
    if spec is None or not spec.submodule_search_locations:
        pytest.skip("vLLM not installed")
    path = os.path.join(list(spec.submodule_search_locations)[0], "distributed", "device_communicators", "pynccl_wrapper.py")
    if not os.path.exists(path):
        pytest.skip("this vLLM has no pynccl_wrapper.py")
    bound = set(re.findall(r'Function\(\s*"(nccl[A-Za-z0-9_]+)"', open(path).read()))
    assert len(bound) >= 10, bound
    _build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", str(_build.nccl_shim_path())], capture_output=True, text=True).stdout
    ours = set(re.findall(r"\bnccl[A-Za-z0-9_]+", out))
    assert not (bound - ours), f"vLLM binds {sorted(bound - ours)}"
