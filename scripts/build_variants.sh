#!/bin/bash
# builds kernel-configuration variants of librst_align.so into _lib/variants/ (experiments only)
# usage: build_variants.sh <<< "name flags..." (one variant per line)
set -e
cd "$(dirname "$0")/.."
mkdir -p realsensetracker_b200/_lib/variants
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC -cudart static -Xptxas -v -I include -ccbin /usr/bin/g++"
while read -r name flags; do
  [ -z "$name" ] && continue
  $NV $flags -o realsensetracker_b200/_lib/variants/$name.so realsensetracker_b200/csrc/*.cu > realsensetracker_b200/_lib/variants/$name.log 2>&1
  echo "$name: $(grep -A2 'k_icp_iterILi0ELb0ELb0' realsensetracker_b200/_lib/variants/$name.log | grep -E 'registers|spill' | tr '\n' ' ' | sed 's/ptxas info    ://; s/  */ /g')"
done
