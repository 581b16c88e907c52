#!/bin/bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
# Builds oracle/_ref/ from the reference sources WHERE THEY LIE (read-only checkout):
#   libtfops_ref_cpu.so : the reference's CPU functions (sliced by line range, see ref_shim.cpp)
#   libtfops_ref_gpu.so : the reference's unmodified .cu kernels compiled for sm_100a
#                         ("the kernel to beat"; launchers keep their C++-mangled names)
#   evaluate_object_3d_offline(_low_iou) : the reference's KITTI AP evaluator, compiled from its own source; the
#                         boost::geometry / ublas names it uses (boost is not installed) come from oracle/boost_shim
#                         (convex-polygon clipping) -- everything else of the program is the reference's
# Nothing from the reference is copied into tracked files; oracle/_ref/ is git-ignored.
# When /root/reference is absent (GPU box) this is a no-op and the prebuilt files are used.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${MONOPSR_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/tf_ops" ]; then
  echo "build_ref: $REF not present; keeping prebuilt oracle/_ref" ; exit 0
fi
mkdir -p "$OUT"
sed -n '21,43p'  "$REF/src/tf_ops/nn_distance/tf_nndistance.cpp"  > "$OUT/nnsearch.inc"
sed -n '23,140p' "$REF/src/tf_ops/approxmatch/tf_approxmatch.cpp" > "$OUT/approxmatch.inc"
# same flags as tf_*_compile.sh:11 (g++ -O2, no -march, no fast-math); contraction off = no FMA
g++ -std=c++11 -O2 -ffp-contract=off -fPIC -shared -I"$HERE" "$HERE/ref_shim.cpp" -o "$OUT/libtfops_ref_cpu.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
if [ -x "$NVCC" ]; then
  "$NVCC" -O2 -gencode arch=compute_100a,code=sm_100a -DGOOGLE_CUDA=1 -Xcompiler -fPIC -shared \
     "$REF/src/tf_ops/nn_distance/tf_nndistance_g.cu" "$REF/src/tf_ops/approxmatch/tf_approxmatch_g.cu" \
     -o "$OUT/libtfops_ref_gpu.so"
fi
EVAL="$REF/scripts/offline_eval/kitti_native_eval"
if [ -f "$EVAL/evaluate_object_3d_offline.cpp" ]; then
  for v in evaluate_object_3d_offline evaluate_object_3d_offline_low_iou; do
    g++ -O2 -std=c++11 -w -I"$HERE/boost_shim" -I"$EVAL" "$EVAL/$v.cpp" -o "$OUT/$v"
  done
fi
echo "build_ref: ok -> $OUT"
