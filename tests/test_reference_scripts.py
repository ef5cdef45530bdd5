"""The reference's OWN p2p test scripts (p2p/tests/*.py, written against ``uccl.p2p``) executed unmodified against this
library through the ``uccl`` module aliases -- on a GPU-less machine the endpoint they create for "GPU 0" runs in host
mode (UCCL_B200_P2P_HOST_FALLBACK=1), which is what those scripts exercise anyway: host tensors, the TCP control plane,
send / recv matching, registration, endpoint removal.  Skipped where the reference tree is not mounted."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference/p2p/tests"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPTS = [
    ("test_engine_metadata.py", "All UCCL P2P Engine tests completed"),
    ("test_engine_send.py", "All UCCL P2P Engine tests completed"),
    ("test_remove_remote_endpoint.py", "0 failed"),
    ("test_gpu_index_mapping.py", "5/5 passed"),
    ("test_register_memory_cache.py", None),
    ("test_util_interval_tree.py", "Exact matches: 2 intervals"),
]


@pytest.mark.parametrize("script,marker", SCRIPTS)
def test_reference_p2p_script(script, marker):
    path = os.path.join(REF, script)
    if not os.path.exists(path):
        pytest.skip("reference tree not available")
    env = dict(os.environ, UCCL_B200_P2P_HOST_FALLBACK="1", UCCL_B200_ALIAS_TOPLEVEL_UTILS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_script.py"), path], capture_output=True,
                       text=True, timeout=300, env=env, cwd=ROOT)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert "Traceback" not in out and "AssertionError" not in out, out[-3000:]
    if marker:
        assert marker in out, out[-3000:]
