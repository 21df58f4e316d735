#!/bin/bash
# Compile the native integrator plugin (native/b200_path_native.cpp) against the headers of the reference checkout and the
# build tree of oracle/build_ref.sh, and put it next to the reference's own plugins in oracle/_ref (git-ignored).
#   native/build_shim.sh [BUILD_DIR]          default BUILD_DIR = /tmp/mi_ref_build
# One g++ call on OUR source file; flags and include paths are those the reference's build uses for src/integrators/path.cpp
# (BUILD_DIR/build.ninja). Test infrastructure: nothing under mitsuba3_b200/ depends on it.
set -e
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(dirname "$HERE")
B=${1:-/tmp/mi_ref_build}; R=/root/reference
[ -f "$B/libmitsuba.so" ] || { echo "no reference build tree at $B (run oracle/build_ref.sh)"; exit 1; }
DST="$ROOT/oracle/_ref/mitsuba_build/plugins"
[ -d "$DST" ] || { echo "no runtime snapshot at $DST"; exit 1; }
/usr/bin/g++ -O2 -DNDEBUG -std=gnu++17 -fPIC -fvisibility=default -march=native -fno-math-errno -fno-trapping-math -fno-strict-aliasing \
  -DLITTLE_ENDIAN -DMI_ENABLE_AUTODIFF=1 -DMI_ENABLE_EMBREE=1 -DMI_ENABLE_LLVM=1 \
  -I"$ROOT/include" -I$R/include -I$R/ext/tinyformat -I$R/ext/nanobind/include -I"$B/include" -I$R/ext/embree/include \
  -I$R/ext/nanobind/ext/robin_map/include -I$R/ext/struct-jit/include -I$R/ext/drjit/include -I$R/ext/drjit/ext/drjit-core/include \
  -I$R/ext/drjit/ext/drjit-core/ext/lz4 -I$R/ext/drjit/ext/drjit-core/ext/nanothread/include \
  -shared "$HERE/b200_path_native.cpp" -o "$DST/b200_path_native.so" \
  -L"$B" -lmitsuba -ldrjit-core -ldrjit-extra -lnanothread -ldl -Wl,-rpath,'$ORIGIN/..'
echo "built $DST/b200_path_native.so"
