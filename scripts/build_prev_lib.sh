#!/bin/bash
# Build the native library of an EARLIER commit next to the current one, for same-box A/B runs of two builds:
#   scripts/build_prev_lib.sh <commit>     ->  deep-learning-for-surgical-video-analysis_b200/lib/libsurgvid_prev.so
#   SURGVID_LIB=$PWD/deep-learning-for-surgical-video-analysis_b200/lib/libsurgvid_prev.so python bench.py ...   (see scripts/gpu_call25.sh)
# profiles/r02/ab*_build_*.json were taken with <commit> = f734fa4 (the build before the elect.sync / CTA-pair work).
set -e
C=${1:?usage: build_prev_lib.sh <commit>}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG=deep-learning-for-surgical-video-analysis_b200
T=$(mktemp -d)
git -C "$ROOT" archive "$C" $PKG/csrc include | tar -x -C "$T"
cd "$T/$PKG/csrc"; mkdir -p build
for f in *.cu; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I"$T/include" -c "$f" -o "build/${f%.cu}.o" & done; wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/$PKG/lib/libsurgvid_prev.so" build/*.o
rm -rf "$T"; ls -la "$ROOT/$PKG/lib/"
