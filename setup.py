"""Builds the native core in-tree (nvcc -gencode arch=compute_100a,code=sm_100a) before packaging.
`pip install -e .` / `python setup.py build_ext --inplace` both end up in uccl_b200/_build.py, the same
entry point `__graft_entry__.build()` uses."""
import importlib.util
import pathlib

from setuptools import setup
from setuptools.command.build_py import build_py

ROOT = pathlib.Path(__file__).parent


def _native_build():
    spec = importlib.util.spec_from_file_location("_ub_build", ROOT / "uccl_b200" / "_build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()


class BuildWithNative(build_py):
    def run(self):
        _native_build()
        super().run()


setup(cmdclass={"build_py": BuildWithNative})
