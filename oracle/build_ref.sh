#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference (sources stay under $REF, default /root/reference)
# into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).  Nothing is copied from $REF.
#
#   ref_cpu_f32 / ref_cpu_f64     g++ only: reference CPU sources + oracle/ref_harness.cpp
#   ref_gpu_f32 / ref_gpu_f64     nvcc sm_100: same + the reference's two .cu files (time stepping)
#   MF_LBM_CUDA_f32 / _f64        the stock program (src/main.cpp), flags of the reference Makefile:60-66
#   MF_LBM_CUDA_shim_f32 / _f64   the reference's src/main.cpp + CPU sources, unmodified, with integration/mflbm_shim.cpp IN
#                                 PLACE OF its two .cu files and linked against libmflbm.so (g++ only): the drop-in, compiled
#
# Precision: the reference hard-codes PRECISION in includes/solver_precision.h:8; oracle/shim/force_f*.h is
# force-included first and pre-defines that header's include guard (see the shim for details).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
CUDA_INC="${CUDA_INC:-/usr/local/cuda/include}"
if [ ! -d "$REF/src" ]; then echo "build_ref: $REF not present - keeping prebuilt files in $OUT"; exit 0; fi
mkdir -p "$OUT"
CPU_SRCS=""
for f in Geometry_preprocessing IO_multiphase Init_multiphase Misc Monitor Phase_gradient utils; do CPU_SRCS="$CPU_SRCS $REF/src/$f.cpp"; done
WHAT="${1:-all}"
for P in f32 f64; do
  SHIM="$HERE/shim/force_$P.h"
  if [ "$WHAT" = all ] || [ "$WHAT" = cpu ]; then
    g++ -std=c++17 -O3 -w -DREF_NO_GPU -include "$SHIM" -I "$REF/includes" -I "$CUDA_INC" \
        $CPU_SRCS "$HERE/ref_harness.cpp" -o "$OUT/ref_cpu_$P" -lm &
  fi
  if [ "$WHAT" = all ] || [ "$WHAT" = gpu ]; then
    NVFLAGS="-gencode arch=compute_100,code=sm_100 -std=c++17 -rdc=true -O3 -lineinfo -w -include $SHIM -I $REF/includes"
    nvcc $NVFLAGS $REF/src/main_iteration_GPU.cu $REF/src/Init_multiphase_GPU.cu $CPU_SRCS "$HERE/ref_harness.cpp" \
         -o "$OUT/ref_gpu_$P" &
    nvcc $NVFLAGS $REF/src/main_iteration_GPU.cu $REF/src/Init_multiphase_GPU.cu $CPU_SRCS $REF/src/main.cpp \
         -o "$OUT/MF_LBM_CUDA_$P" &
  fi
  if [ "$WHAT" = all ] || [ "$WHAT" = shim ]; then
    LIBDIR="$HERE/../mf-lbm-cuda_b200/lib"
    if [ -f "$LIBDIR/libmflbm.so" ]; then
      g++ -std=c++17 -O3 -w -include "$SHIM" -I "$REF/includes" -I "$CUDA_INC" -I "$HERE/../include" \
          $CPU_SRCS "$REF/src/main.cpp" "$HERE/../integration/mflbm_shim.cpp" -o "$OUT/MF_LBM_CUDA_shim_$P" \
          -L "$LIBDIR" -lmflbm -Wl,-rpath,'$ORIGIN/../../mf-lbm-cuda_b200/lib' -lm &
    else
      echo "build_ref: $LIBDIR/libmflbm.so missing - shim programs skipped"
    fi
  fi
done
wait
ls -la "$OUT"
