#!/bin/bash
# Compiles the reference's own per-element code (headers under /root/reference, never copied)
# into oracle/_ref/libitmref.so. Outputs go only to oracle/_ref/ (git-ignored, ships via gpurun).
# Flags mirror a host-only ITMLib build: COMPILE_WITHOUT_CUDA, and __device__ stubbed because
# buildHashAllocAndVisibleTypePP is declared device-only (DA/ITMSceneReconstructionEngine.h:176).
# No contraction / fast-math so the reference code is evaluated in plain IEEE binary32.
set -e
HERE_EARLY=1
REF=${REF:-/root/reference/src/InfiniTAM/InfiniTAM}
HERE="$(cd "$(dirname "$0")" && pwd)"
if [ ! -d "$REF/ITMLib" ]; then echo "reference not present at $REF; keeping prebuilt oracle/_ref" >&2; exit 0; fi
mkdir -p "$HERE/_ref"
/usr/bin/g++ -std=c++14 -O2 -ffp-contract=off -fno-fast-math -shared -fPIC -w \
    -DCOMPILE_WITHOUT_CUDA -D__device__= -I"$REF" \
    -o "$HERE/_ref/libitmref.so" "$HERE/ref_driver.cpp" "$REF/ITMLib/Engine/DeviceSpecific/CPU/ITMMeshingEngine_CPU.cpp"
echo "built $HERE/_ref/libitmref.so"

# ---- instance frame splitting / compositing: the reference's own free functions (oracle/_ref/libinstrecref.so) ---------
# InstanceReconstructor.cpp needs OpenCV/Eigen/Pangolin as a translation unit, but ProcessSilhouette_CPU, RemoveSilhouette_CPU,
# CompositeDepth and CompositeColor only need ORUtils images, the reference's Mask/BoundingBox headers, a byte matrix and two
# integer vectors (oracle/stubs, oracle/ref_frames_driver.cpp). They are cut out of the reference file here, at build time,
# into oracle/_ref/ (git-ignored) and compiled unmodified.
DS=${DS:-/root/reference/src/DynSLAM}
IRC="$DS/InstRecLib/InstanceReconstructor.cpp"
if [ -f "$IRC" ]; then
  awk '/^template <typename DEPTH_T>/ {on=1} /^void InstanceReconstructor::ProcessFrame\(/ {on=0} on' "$IRC" > "$HERE/_ref/instrec_extract.inc"
  awk '/^void CompositeDepth\(/ {on=1} /^void InstanceReconstructor::CompositeInstanceDepthMaps\(/ {on=0} on' "$IRC" >> "$HERE/_ref/instrec_extract.inc"
  if grep -q "ProcessSilhouette_CPU" "$HERE/_ref/instrec_extract.inc" && grep -q "void CompositeColor" "$HERE/_ref/instrec_extract.inc"; then
    /usr/bin/g++ -std=c++14 -O2 -ffp-contract=off -fno-fast-math -shared -fPIC -w -DCOMPILE_WITHOUT_CUDA -D__device__= \
        -I"$HERE/stubs" -I"$REF" -I"$DS/InstRecLib/Utils" -I"$HERE" \
        -o "$HERE/_ref/libinstrecref.so" "$HERE/ref_frames_driver.cpp"
    echo "built $HERE/_ref/libinstrecref.so"
    rm -f "$HERE/_ref/instrec_extract.inc"   # the cut-out text is a build intermediate only
  else
    echo "could not locate the silhouette / compositing functions in $IRC" >&2; rm -f "$HERE/_ref/instrec_extract.inc"
  fi
fi

# ---- evaluation consumer: the reference's own ProjectLidar / EvaluateDepth / ComputeAccuracy (oracle/_ref/libevalref.so) ---------
# Evaluation.cpp and EvaluationCallback.cpp need Eigen, OpenCV, Pangolin and the DynSlam class as translation units; the four
# function bodies only need what oracle/ref_eval_driver.cpp declares around them. They are cut out of the reference files here,
# at build time, into oracle/_ref/ (git-ignored) and compiled unmodified.
EVC="$DS/Evaluation/Evaluation.cpp"; ECB="$DS/Evaluation/EvaluationCallback.cpp"
if [ -f "$EVC" ] && [ -f "$ECB" ]; then
  awk '/^bool Evaluation::ProjectLidar\(/ {on=1} /^\/\/ Track = ours, tracklet = ground truth/ {on=0} on' "$EVC" > "$HERE/_ref/eval_extract.inc"
  awk '/^void EvaluationCallback::ProcessLidarPoint\(/ {on=1} /^DepthEvaluation EvaluationCallback::CreateDepthEvaluation\(/ {on=0} on' "$ECB" >> "$HERE/_ref/eval_extract.inc"
  awk '/^void EvaluationCallback::ComputeAccuracy\(/ {on=1} on {print} on && /^}$/ {on=0}' "$ECB" >> "$HERE/_ref/eval_extract.inc"
  if grep -q "Evaluation::EvaluateDepth" "$HERE/_ref/eval_extract.inc" && grep -q "EvaluationCallback::ComputeAccuracy" "$HERE/_ref/eval_extract.inc"; then
    /usr/bin/g++ -std=c++14 -O2 -ffp-contract=off -fno-fast-math -shared -fPIC -w -I"$HERE/stubs" -I"$DS/Evaluation" -I"$HERE" \
        -o "$HERE/_ref/libevalref.so" "$HERE/ref_eval_driver.cpp" && echo "built $HERE/_ref/libevalref.so"
    rm -f "$HERE/_ref/eval_extract.inc"   # the cut-out text is a build intermediate only
  else
    echo "could not locate the evaluation functions in $EVC / $ECB" >&2; rm -f "$HERE/_ref/eval_extract.inc"
  fi
fi

# ---- on-disk formats: the reference's own ReadFilePFM, ReadMask and ITMMesh::WriteOBJ (oracle/_ref/libioref.so) -----------------
PFM=${PFM:-/root/reference/src/pfmLib/ImageIOpfm.cpp}; PSP="$DS/InstRecLib/PrecomputedSegmentationProvider.cpp"
if [ -f "$PFM" ] && [ -f "$PSP" ]; then
  awk '/^int ReadFilePFM\(/ {inr=1} {print} inr && /^}$/ {exit}' "$PFM" > "$HERE/_ref/pfm_extract.inc"
  awk '/^uint8_t \*ReadMask\(/ {on=1} on {print} on && /^}$/ {exit}' "$PSP" > "$HERE/_ref/mask_extract.inc"
  if grep -q "int ReadFilePFM" "$HERE/_ref/pfm_extract.inc" && grep -q "ReadMask" "$HERE/_ref/mask_extract.inc"; then
    /usr/bin/g++ -std=c++14 -O2 -shared -fPIC -w -DCOMPILE_WITHOUT_CUDA -I"$HERE/stubs" -I"$REF" -I"$(dirname "$PFM")" -I"$HERE" \
        -o "$HERE/_ref/libioref.so" "$HERE/ref_io_driver.cpp" && echo "built $HERE/_ref/libioref.so"
  else
    echo "could not locate ReadFilePFM / ReadMask" >&2
  fi
  rm -f "$HERE/_ref/pfm_extract.inc" "$HERE/_ref/mask_extract.inc"   # the cut-out text is a build intermediate only
fi

# ---- reference CUDA build + ITMLib harness (oracle/_ref/libitmharness.so) -------------------------
# The reference's own CUDA engines, unmodified, compiled per-TU for sm_100a with the reference's
# flags (--use_fast_math, ITMLib/CMakeLists.txt:226-230) directly from /root/reference, plus the few
# host TUs they need, plus oracle/itm_harness.cpp which drives them (and the B200 shim) through the
# real ITMLib interfaces. libb200fusion.so must already be built (python __graft_entry__.py).
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
LIBB200="$HERE/../dynslam_b200/csrc/libb200fusion.so"
if [ ! -f "$LIBB200" ]; then echo "libb200fusion.so missing; skipping harness" >&2; exit 0; fi
OBJ="$HERE/_ref/obj"; mkdir -p "$OBJ"
CU="ITMLib/Engine/DeviceSpecific/CUDA/ITMSceneReconstructionEngine_CUDA.cu ITMLib/Engine/DeviceSpecific/CUDA/ITMVisualisationEngine_CUDA.cu ITMLib/Engine/DeviceSpecific/CUDA/ITMMeshingEngine_CUDA.cu ITMLib/Engine/DeviceSpecific/CUDA/ITMViewBuilder_CUDA.cu"
CPP="ITMLib/Utils/ITMLibSettings.cpp ITMLib/Objects/ITMPose.cpp ITMLib/Engine/ITMVisualisationEngine.cpp ORUtils/CUDADefines.cpp"
pids=""
for f in $CU; do
  o="$OBJ/$(basename $f .cu).o"
  if [ ! -f "$o" ] || [ "$REF/$f" -nt "$o" ]; then
    $NVCC -std=c++14 -gencode arch=compute_100a,code=sm_100a --use_fast_math -O3 -w -Xcompiler -fPIC -ccbin /usr/bin/g++ -I"$REF" -c "$REF/$f" -o "$o" &
    pids="$pids $!"
  fi
done
for f in $CPP; do
  o="$OBJ/$(basename $f .cpp).o"
  if [ ! -f "$o" ]; then /usr/bin/g++ -std=c++14 -O2 -w -fPIC -I"$REF" -I/usr/local/cuda/include -c "$REF/$f" -o "$o" & pids="$pids $!"; fi
done
for p in $pids; do wait $p; done
/usr/bin/g++ -std=c++14 -O2 -w -fPIC -I"$REF" -I/usr/local/cuda/include -I"$HERE/../include" -I"$HERE/../dynslam_b200/itm_shim" \
    -c "$HERE/itm_harness.cpp" -o "$OBJ/itm_harness.o"
$NVCC -shared -o "$HERE/_ref/libitmharness.so" "$OBJ"/*.o -ccbin /usr/bin/g++ -Xlinker -rpath -Xlinker '$ORIGIN/../../dynslam_b200/csrc' \
    -L"$HERE/../dynslam_b200/csrc" -lb200fusion -lcudart 2>&1 | grep -v "deprecated" || true
echo "built $HERE/_ref/libitmharness.so"
