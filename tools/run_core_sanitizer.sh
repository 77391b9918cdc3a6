#!/bin/bash
# ASAN + UBSAN build of the C++ scheduler core (csrc/alloc: allocator / exact solver / stimulator)
# exercised by the allocator and stimulator CPU tests (SURVEY §4 "sanitizers").  The instrumented
# module is built into a scratch directory and put in front of the package's own _core through
# SKY_CORE_OVERRIDE, so the in-tree .so is never replaced.
set -euo pipefail
cd "$(dirname "$0")/.."
OUT=${1:-/tmp/sky_core_asan}
mkdir -p "$OUT"
PYINC=$(python - <<'PY'
import sysconfig, pybind11
print("-I" + sysconfig.get_paths()["include"], "-I" + pybind11.get_include())
PY
)
EXT=$(python -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
g++ -O1 -g -std=c++17 -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer $PYINC -Icsrc \
    -shared csrc/alloc/allocator.cc csrc/alloc/stimulator.cc csrc/bindings_core.cpp \
    -o "$OUT/_core$EXT" -lpthread
ASAN_LIB=$(g++ -print-file-name=libasan.so)
# libstdc++ must be loaded before ASAN resolves its __cxa_throw interceptor (python itself does
# not link it), otherwise the first C++ exception aborts inside the sanitizer runtime
STDCXX=$(g++ -print-file-name=libstdc++.so.6)
LD_PRELOAD="$ASAN_LIB $STDCXX" ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 UBSAN_OPTIONS=halt_on_error=1 \
  SKY_CORE_OVERRIDE="$OUT" python -m pytest tests/test_allocator_cpu.py \
  "tests/test_core_cpu.py::test_stimulator_matches_numpy_reference_streams" -x -q -p no:cacheprovider
