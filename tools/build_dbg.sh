#!/bin/bash
# Debug build of the library with the epilogue phase timers (clock64 + printf) compiled in:
#   tools/build_dbg.sh && ISWM_B200_LIB=$PWD/gpurun_out/libiswm_b200_dbg.so python tools/prof_conv.py ...
set -e
cd "$(dirname "$0")/../iswm_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -DISWM_EPI_TIMING"
mkdir -p /tmp/iswm_dbg ../../gpurun_out
for f in lib loss_metric elementwise bn_dual stem_pool sgd_pack tc_host conv_igemm conv_wgrad peer_allreduce; do $NVCC $FLAGS -c $f.cu -o /tmp/iswm_dbg/$f.o & done; wait
$NVCC -shared -o ../../tools/libiswm_b200_dbg.so /tmp/iswm_dbg/*.o -gencode arch=compute_100a,code=sm_100a
echo built tools/libiswm_b200_dbg.so
