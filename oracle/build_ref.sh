#!/bin/bash
# Build the UNMODIFIED reference from the sources where they lie (/root/reference, read-only)
# and snapshot its runtime into oracle/_ref/mitsuba_build (git-ignored, travels with gpurun).
#
#   oracle/build_ref.sh [--embree on|off] [--jobs N] [--build-dir DIR]
#
# Recipe = SURVEY.md Appendix A with MI_ENABLE_EMBREE=ON (/root/reference/CMakeLists.txt:26),
# variants scalar_rgb + llvm_ad_rgb.  The intermediate tree lives under /tmp (several GB of
# objects; not product, not shipped); only the runtime files (libmitsuba, plugins, python
# packages, data) are copied into oracle/_ref.  Test infrastructure: used by
# tests/golden/gen_golden*.py (fixtures), tests/test_mitsuba_plugin.py (live plugin) and
# bench.py's reference arm.  Nothing under mitsuba3_b200/ depends on it.
#
# `--embree off` selects the reference's own kd-tree + Moeller-Trumbore (mesh.h:1132-1153):
# that is the build the bit-exact hit fixtures (tests/golden/cbox_*.npz) were generated with.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
EMBREE=ON; JOBS=$(nproc); BUILD=/tmp/mi_ref_build
while [ $# -gt 0 ]; do
  case "$1" in
    --embree) [ "$2" = off ] && EMBREE=OFF || EMBREE=ON; shift 2;;
    --jobs) JOBS=$2; shift 2;;
    --build-dir) BUILD=$2; shift 2;;
    *) echo "unknown argument $1"; exit 2;;
  esac
done
[ -f /root/reference/CMakeLists.txt ] || { echo "no reference checkout at /root/reference"; exit 1; }
PY=$(command -v python3)
cmake -S /root/reference -B "$BUILD" -GNinja -DCMAKE_BUILD_TYPE=Release \
      -DCMAKE_C_COMPILER=/usr/bin/gcc -DCMAKE_CXX_COMPILER=/usr/bin/g++ \
      -DPython_EXECUTABLE="$PY" \
      -DMI_DEFAULT_VARIANTS="scalar_rgb,llvm_ad_rgb" -DMI_ENABLE_EMBREE=$EMBREE
ninja -C "$BUILD" -j"$JOBS"
echo "$EMBREE" > "$BUILD/EMBREE_SETTING"
"$HERE/ref_snapshot.sh" "$BUILD"
# LLVM for llvm_ad_rgb: the image has no libLLVM.so; oracle/llvm_shim builds one over the
# LLVM that llvmlite carries (see oracle/llvm_shim/README.md).
if [ -f "$HERE/llvm_shim/Makefile" ]; then make -C "$HERE/llvm_shim" || echo "llvm shim not built (llvm_ad_rgb unavailable)"; fi
# the compiled native integrator plugin (native/b200_path_native.cpp) goes next to the reference's own plugins
if [ -x "$HERE/../native/build_shim.sh" ]; then "$HERE/../native/build_shim.sh" "$BUILD" || echo "native shim not built"; fi
