#!/bin/bash
# Build libsfxm_b200.so for sm_100a (B200).  -fmad=false: the parity target is the reference's plain IEEE
# multiply/add order; fused multiply-adds would change the last bit of every residual.
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC ${SFXM_NVCC_EXTRA}"
cd "$HERE"
# Nothing to do when the library is newer than every source and header (the GPU boxes receive the built .so but not
# the object files, see .gpurunignore: a rebuild there would cost minutes of GPU time).
if [ -f libsfxm_b200.so ] && [ -z "${SFXM_NVCC_EXTRA}" ]; then
  stale=0
  for src in *.cu *.cuh *.h ../../include/sfxm_b200.h; do
    if [ "$src" -nt libsfxm_b200.so ]; then stale=1; break; fi
  done
  if [ "$stale" = 0 ]; then
    echo "up to date: $HERE/libsfxm_b200.so"
    exit 0
  fi
fi
SRCS="sfxm_host sfxm_dispatch sfxm_inst_ad1 sfxm_inst_ad2 sfxm_inst_ad3 sfxm_inst_ad4 sfxm_inst_euler_prim sfxm_inst_euler_cons sfxm_fused_ad3 sfxm_fused_euler sfxm_solve sfxm_newton_ctl sfxm_newton_driver sfxm_ode sfxm_selftest sfxm_amr sfxm_pcr sfxm_pcr4 sfxm_spike sfxm_stencil_api"
for f in $SRCS; do
  if [ ! -f "$f.o" ] || [ "$f.cu" -nt "$f.o" ] || [ sfxm_device.cuh -nt "$f.o" ] || [ sfxm_kernels.cuh -nt "$f.o" ] || [ sfxm_reduce.cuh -nt "$f.o" ] || [ sfxm_pcr.cuh -nt "$f.o" ] || [ sfxm_thomas.cuh -nt "$f.o" ] || [ sfxm_cpasync.cuh -nt "$f.o" ] || [ sfxm_fused.cuh -nt "$f.o" ] || [ ../../include/sfxm_b200.h -nt "$f.o" ] || [ sfxm_host.h -nt "$f.o" ]; then
    "$NVCC" $FLAGS -c "$f.cu" -o "$f.o" &
  fi
done
wait
OBJS=""; for f in $SRCS; do OBJS="$OBJS $f.o"; done
"$NVCC" -shared -o libsfxm_b200.so $OBJS -lcudart
echo "built $HERE/libsfxm_b200.so"
