#!/bin/bash
# builds kernel-configuration variants of librst_align.so into _lib/variants/ (experiments only)
set -e
cd "$(dirname "$0")/.."
mkdir -p realsensetracker_b200/_lib/variants
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC -cudart static -Xptxas -v -I include -ccbin /usr/bin/g++"
while read -r name flags; do
  [ -z "$name" ] && continue
  $NV $flags -o realsensetracker_b200/_lib/variants/$name.so realsensetracker_b200/csrc/*.cu > realsensetracker_b200/_lib/variants/$name.log 2>&1
  echo "$name: $(grep -A2 'k_icp_iterILi0ELb0ELb0' realsensetracker_b200/_lib/variants/$name.log | grep -E 'registers|spill' | tr '\n' ' ' | sed 's/ptxas info    ://; s/  */ /g')"
done <<VARS
v_t256_c4_b2 -DRST_ICP_THREADS=256 -DRST_ICP_CPW=4 -DRST_ICP_MINB=2
v_t256_c2_b3 -DRST_ICP_THREADS=256 -DRST_ICP_CPW=2 -DRST_ICP_MINB=3
v_t256_c2_b2 -DRST_ICP_THREADS=256 -DRST_ICP_CPW=2 -DRST_ICP_MINB=2
v_t128_c2_b5 -DRST_ICP_THREADS=128 -DRST_ICP_CPW=2 -DRST_ICP_MINB=5
v_t128_c2_b6 -DRST_ICP_THREADS=128 -DRST_ICP_CPW=2 -DRST_ICP_MINB=6
v_t128_c4_b4 -DRST_ICP_THREADS=128 -DRST_ICP_CPW=4 -DRST_ICP_MINB=4
v_t128_c1_b8 -DRST_ICP_THREADS=128 -DRST_ICP_CPW=1 -DRST_ICP_MINB=8
VARS
