"""Static drift guards that run without a GPU:

* every keyword argument that the GPU tests, `bench.py`, `__graft_entry__.py`, the benchmarks and the examples pass to
  a function or method of this package exists in some signature of that name (the GPU suite only runs at the end of a
  round: an API rename must not wait until then to be noticed);
* every ``from uccl_b200... import name`` in the fenced python blocks of the documentation resolves.
"""
import ast
import glob
import importlib
import inspect
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MODULES = ["uccl_b200", "uccl_b200.p2p", "uccl_b200.p2p.utils", "uccl_b200.p2p.internode", "uccl_b200.p2p.compress",
           "uccl_b200.collective", "uccl_b200.ep", "uccl_b200.ep.buffer", "uccl_b200.ep.low_latency",
           "uccl_b200.ep.utils", "uccl_b200.ep.host_ep", "uccl_b200.ep.proxy", "uccl_b200.ep.autograd",
           "uccl_b200.ukernel", "uccl_b200.ukernel.dsl", "uccl_b200.ukernel.p2p", "uccl_b200.parallel",
           "uccl_b200.parallel.comm", "uccl_b200.parallel.ddp", "uccl_b200.parallel.multinode", "uccl_b200.parallel.pg",
           "uccl_b200.net", "uccl_b200.net.topology", "uccl_b200.ops", "uccl_b200.utils", "uccl_b200.utils.regions",
           "uccl_b200.utils.sm_partition", "uccl_b200.utils.tuner", "uccl_b200.utils.metrics", "uccl_b200.models",
           "uccl_b200.models.moe", "uccl_b200.models.resnet", "deep_ep", "deep_ep.buffer"]
# names that torch / the standard library also use with other keywords
GENERIC = {"zeros", "empty", "ones", "full", "randn", "rand", "arange", "run", "Event", "view", "sum", "wait", "barrier",
           "synchronize", "clone", "to", "Stream", "stream", "get", "put", "copy_", "add", "join", "open",
           "Thread", "Process", "main", "close"}


def _signatures():
    sigs = {}

    def add(name, fn):
        try:
            sigs.setdefault(name, []).append(inspect.signature(fn))
        except (TypeError, ValueError):
            pass

    for m in MODULES:
        mod = importlib.import_module(m)
        for n, o in vars(mod).items():
            if inspect.isfunction(o):
                add(n, o)
            elif inspect.isclass(o) and getattr(o, "__module__", "").split(".")[0] in ("uccl_b200", "deep_ep"):
                add(n, o)
                for k, v in vars(o).items():
                    f = v.__func__ if isinstance(v, (staticmethod, classmethod)) else v
                    if inspect.isfunction(f):
                        add(k, f)
    return sigs


def test_keyword_arguments_used_by_gpu_tests_and_benches_exist():
    sigs = _signatures()
    files = sorted(set(glob.glob(os.path.join(ROOT, "tests", "*gpu*.py")) + glob.glob(os.path.join(ROOT, "benchmarks", "*.py"))
                       + glob.glob(os.path.join(ROOT, "examples", "*.py"))
                       + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]))
    assert len(files) > 20
    problems = []
    for f in files:
        tree = ast.parse(open(f).read())
        for node in ast.walk(tree):
            if not (isinstance(node, ast.Call) and node.keywords):
                continue
            fn = node.func
            name = fn.attr if isinstance(fn, ast.Attribute) else (fn.id if isinstance(fn, ast.Name) else None)
            if name not in sigs or name in GENERIC:
                continue
            kws = [k.arg for k in node.keywords if k.arg]
            ok = False
            for s in sigs[name]:
                ps = s.parameters
                if any(p.kind == p.VAR_KEYWORD for p in ps.values()) or all(k in ps for k in kws):
                    ok = True
                    break
            if not ok:
                problems.append(f"{os.path.relpath(f, ROOT)}:{node.lineno} {name}({', '.join(kws)})")
    assert not problems, "\n".join(problems)


def test_documentation_snippets_import_existing_names():
    docs = [os.path.join(ROOT, "README.md"), os.path.join(ROOT, "DESIGN.md")] + glob.glob(os.path.join(ROOT, "docs", "*.md"))
    checked, problems = 0, []
    for f in docs:
        for m in re.finditer(r"```python\n(.*?)```", open(f).read(), re.S):
            try:
                tree = ast.parse(m.group(1))
            except SyntaxError as e:
                problems.append(f"{os.path.relpath(f, ROOT)}: snippet does not parse: {e}")
                continue
            for n in ast.walk(tree):
                if isinstance(n, ast.ImportFrom) and n.module and n.module.split(".")[0] in ("uccl_b200", "deep_ep"):
                    mod = importlib.import_module(n.module)
                    for a in n.names:
                        checked += 1
                        if not hasattr(mod, a.name):
                            try:
                                importlib.import_module(n.module + "." + a.name)
                            except ImportError:
                                problems.append(f"{os.path.relpath(f, ROOT)}: {n.module}.{a.name} does not exist")
    assert checked > 5 and not problems, "\n".join(problems)


def test_every_environment_knob_is_documented():
    """Each UCCL_B200_* variable the sources read appears in docs/config.md (families may be abbreviated there as
    `UCCL_B200_X_A` / `_B`)."""

    names = set()
    for pat in ("*.py", "*.cc", "*.h", "*.cu", "*.cuh"):
        for f in glob.glob(os.path.join(ROOT, "uccl_b200", "**", pat), recursive=True) + [os.path.join(ROOT, "bench.py")]:
            try:
                names.update(re.findall(r"UCCL_B200_[A-Z0-9_]+", open(f, errors="ignore").read()))
            except OSError:
                pass
    doc = open(os.path.join(ROOT, "docs", "config.md")).read()
    assert len(names) > 25
    missing = []
    for n in sorted(names):
        if n in doc or n.endswith("_"):
            continue
        parts = n.split("_")
        # abbreviated family member: some suffix "_X_Y" of the name appears in backticks or after a slash
        if any(("`_" + "_".join(parts[i:]) + "`") in doc or ("/ `_" + "_".join(parts[i:])) in doc or ("_" + "_".join(parts[i:]) + "`") in doc
               for i in range(2, len(parts))):
            continue
        missing.append(n)
    assert not missing, missing


def test_every_prefixed_native_parameter_is_documented():
    """Parameters the native code reads through UB_PARAM / param_load (the UCCL_B200_ prefix is added at run time)."""
    names = set()
    for f in glob.glob(os.path.join(ROOT, "uccl_b200", "csrc", "**", "*.*"), recursive=True):
        if f.endswith((".cc", ".h", ".cu", ".cuh")):
            src = open(f, errors="ignore").read()
            names.update(re.findall(r'UB_PARAM\([A-Za-z0-9_]+,\s*"([A-Z0-9_]+)"', src))
            names.update(re.findall(r'param_load(?:_str)?\("([A-Z0-9_]+)"', src))
    names.discard("ENV_SUFFIX")
    doc = open(os.path.join(ROOT, "docs", "config.md")).read()
    assert len(names) > 30
    missing = []
    for n in sorted(names):
        parts = n.split("_")
        if n in doc or any(("_" + "_".join(parts[i:]) + "`") in doc for i in range(1, len(parts))):
            continue
        missing.append(n)
    assert not missing, missing
