#!/usr/bin/env bash
# Format C++/CUDA with clang-format and Python with black (whichever is installed); --check only reports.
set -euo pipefail
cd "$(dirname "$0")"
MODE=${1:-fix}
CPP=$(git ls-files 'uccl_b200/csrc/**' 'tests/cpp/*' 'benchmarks/*.cc' | grep -E '\.(h|cuh|cu|cc)$' || true)
PY=$(git ls-files '*.py')
if command -v clang-format >/dev/null; then
  if [ "$MODE" = "--check" ]; then clang-format --dry-run --Werror $CPP; else clang-format -i $CPP; fi
else
  echo "clang-format not installed: skipping C++/CUDA" >&2
fi
if python -c 'import black' 2>/dev/null; then
  if [ "$MODE" = "--check" ]; then python -m black --check -l 120 $PY; else python -m black -l 120 $PY; fi
else
  echo "black not installed: skipping Python" >&2
fi
