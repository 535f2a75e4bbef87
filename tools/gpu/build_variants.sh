#!/bin/bash
# builds libeigb200_<tag>.so variants that differ in -D flags of gemm_tma.cu (debugging aid)
set -e
cd "$(dirname "$0")/../.."
L=eigensolver_gpu_b200/lib
python -m eigensolver_gpu_b200.build > /dev/null
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -I eigensolver_gpu_b200/csrc"
for v in "v0:-DEIGB_VAR_NODBG" "v1:-DEIGB_VAR_NODBG -DEIGB_VAR_GC" "v3:-DEIGB_VAR_NODBG -DEIGB_VAR_BRANCH"; do
  tag=${v%%:*}; defs=${v#*:}
  /usr/local/cuda/bin/nvcc $F $defs -c eigensolver_gpu_b200/csrc/gemm_tma.cu -o /tmp/gemm_tma_$tag.o
  objs=$(ls $L/*.o | grep -v gemm_tma.o)
  /usr/local/cuda/bin/nvcc -shared -o $L/libeigb200_$tag.so $objs /tmp/gemm_tma_$tag.o -lcudart -ldl -gencode arch=compute_100a,code=sm_100a
  echo built $L/libeigb200_$tag.so
done
