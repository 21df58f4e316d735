#!/bin/bash
# Snapshot the RUNTIME files of an existing build tree of the UNMODIFIED reference into
# oracle/_ref/mitsuba_build (git-ignored; travels to the GPU box with gpurun).
#
# This does not build anything: the reference needs its own cmake build system, generated
# headers, Dr.Jit and nanobind, which a short Makefile cannot reproduce.
# A build tree made by the survey stage (`cmake -DMI_DEFAULT_VARIANTS="scalar_rgb,llvm_ad_rgb"
# -DMI_ENABLE_EMBREE=OFF`) is reused when it is still around. Without it
# nothing is lost but the optional `kind: "reference"` CPU baseline and the live-Mitsuba plugin
# tests (tests/test_mitsuba_plugin.py), which then skip.
set -e
SRC=${1:-/tmp/mi_probe/build}
DST=$(dirname "$0")/_ref/mitsuba_build
[ -f "$SRC/libmitsuba.so" ] || { echo "no reference build tree at $SRC"; exit 1; }
rm -rf "$DST"; mkdir -p "$DST"
cp -a "$SRC"/*.so "$DST"/
cp -a "$SRC"/plugins "$DST"/
mkdir -p "$DST/python"
cp -a "$SRC"/python/mitsuba "$SRC"/python/drjit "$DST/python/"
find "$DST" -name "*.pyi" -delete; rm -rf "$DST"/python/*/mitsuba_stubs "$DST"/python/drjit/stubs 2>/dev/null || true
# data files the plugins need at run time (sRGB upsampling table etc.)
[ -d "$SRC/data" ] && cp -a "$SRC/data" "$DST"/ || true
du -sh "$DST"
