"""Runs a script written against the reference's ``uccl`` package on this library: installs the module aliases
(`uccl_b200.compat`) and executes the script as ``__main__``.  Used by tests/test_reference_scripts.py.

    python tests/run_reference_script.py /path/to/reference/p2p/tests/test_engine_send.py [args...]
"""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uccl_b200.compat  # noqa: E402

uccl_b200.compat.install()
sys.argv = sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
