#!/bin/bash
# Compiles the reference's own per-element code (headers under /root/reference, never copied)
# into oracle/_ref/libitmref.so. Outputs go only to oracle/_ref/ (git-ignored, ships via gpurun).
# Flags mirror a host-only ITMLib build: COMPILE_WITHOUT_CUDA, and __device__ stubbed because
# buildHashAllocAndVisibleTypePP is declared device-only (DA/ITMSceneReconstructionEngine.h:176).
# No contraction / fast-math so the reference code is evaluated in plain IEEE binary32.
set -e
REF=${REF:-/root/reference/src/InfiniTAM/InfiniTAM}
HERE="$(cd "$(dirname "$0")" && pwd)"
if [ ! -d "$REF/ITMLib" ]; then echo "reference not present at $REF; keeping prebuilt oracle/_ref" >&2; exit 0; fi
mkdir -p "$HERE/_ref"
/usr/bin/g++ -std=c++14 -O2 -ffp-contract=off -fno-fast-math -shared -fPIC -w \
    -DCOMPILE_WITHOUT_CUDA -D__device__= -I"$REF" \
    -o "$HERE/_ref/libitmref.so" "$HERE/ref_driver.cpp"
echo "built $HERE/_ref/libitmref.so"
