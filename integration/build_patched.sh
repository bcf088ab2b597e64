#!/bin/bash
# Applies integration/itmlib_b200.patch to the reference's ITMLib — in a scratch tree of symlinks under oracle/_ref/patched
# (git-ignored; nothing of the reference is copied into the repository) — and builds the WHOLE of ITMLib from it, patched
# constructors of ITMDenseMapper / ITMMainEngine included, into oracle/_ref/libitmpatched.so together with
# integration/main_engine_driver.cpp: ITMMainEngine::ProcessFrame with settings->engineBackend = BACKEND_B200 (or
# BACKEND_REFERENCE) on synthetic frames. This is the binding a DynSLAM maintainer would add, compiled and run
# (tests/test_gpu_itm_harness.py::test_patched_main_engine*). The reference's own build system is not used.
set -e
REF=${REF:-/root/reference/src/InfiniTAM/InfiniTAM}
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$HERE/.."
OUT="$ROOT/oracle/_ref"
if [ ! -d "$REF/ITMLib" ]; then echo "reference not present at $REF; keeping prebuilt oracle/_ref" >&2; exit 0; fi
if [ ! -f "$ROOT/dynslam_b200/csrc/libb200fusion.so" ]; then echo "libb200fusion.so missing (python __graft_entry__.py first)" >&2; exit 0; fi
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
P="$OUT/patched"
rm -rf "$P"; mkdir -p "$P" "$OUT/obj_patched"
cp -rs "$REF/ITMLib" "$P/ITMLib"          # a tree of symlinks: quoted relative includes resolve inside it
cp -rs "$REF/ORUtils" "$P/ORUtils"
for f in ITMLib/Utils/ITMLibSettings.h ITMLib/Utils/ITMLibSettings.cpp ITMLib/Engine/ITMDenseMapper.cpp ITMLib/Engine/ITMMainEngine.cpp; do
  rm "$P/$f"; cp "$REF/$f" "$P/$f"        # build intermediates of the patch (deleted with the tree)
done
(cd "$P" && patch -p1 --no-backup-if-mismatch < "$HERE/itmlib_b200.patch")
INC="-I$P -I$ROOT/include -I$ROOT/dynslam_b200/itm_shim -I/usr/local/cuda/include"
O="$OUT/obj_patched"
pids=""
for f in $(cd "$P" && ls ITMLib/Engine/DeviceSpecific/CUDA/*.cu); do
  ref_obj="$OUT/obj/$(basename $f .cu).o"
  if [ -f "$ref_obj" ]; then ln -sf "$ref_obj" "$O/$(basename $f .cu).o"; continue; fi     # unpatched kernels: objects of build_ref.sh
  $NVCC -std=c++14 -gencode arch=compute_100a,code=sm_100a --use_fast_math -O3 -w -Xcompiler -fPIC -ccbin /usr/bin/g++ -I"$P" -c "$P/$f" -o "$O/$(basename $f .cu).o" &
  pids="$pids $!"
done
for f in $(cd "$P" && ls ITMLib/Engine/*.cpp ITMLib/Engine/DeviceSpecific/CPU/*.cpp ITMLib/Objects/*.cpp ITMLib/Utils/*.cpp ORUtils/*.cpp | grep -v ITMOxtsIO); do
  /usr/bin/g++ -std=c++14 -O2 -w -fPIC $INC -c "$P/$f" -o "$O/$(echo $f | tr '/' '_' | sed 's/\.cpp$/.o/')" &
  pids="$pids $!"
done
/usr/bin/g++ -std=c++14 -O2 -w -fPIC $INC -c "$HERE/main_engine_driver.cpp" -o "$O/main_engine_driver.o" &
pids="$pids $!"
fail=0
for p in $pids; do wait $p || fail=1; done
[ $fail = 0 ] || { echo "compile of the patched ITMLib failed" >&2; exit 1; }
$NVCC -shared -o "$OUT/libitmpatched.so" "$O"/*.o -ccbin /usr/bin/g++ -Xlinker -rpath -Xlinker '$ORIGIN/../../dynslam_b200/csrc' \
    -L"$ROOT/dynslam_b200/csrc" -lb200fusion -lcudart 2>&1 | grep -v "deprecated" || true
rm -rf "$P"
echo "built $OUT/libitmpatched.so"
