#!/bin/bash
# Build lab variants of libapap_b200.so (cross-compiles here): tools/variants.sh name "-DFLAG=1 ..." [name2 "..."]...
# -> cvx_proj_b200/lab/<name>.so ; run on the GPU box with APAP_B200_LIB=cvx_proj_b200/lab/<name>.so
set -e
cd "$(dirname "$0")/../cvx_proj_b200/csrc"
mkdir -p ../lab
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  make -s BUILD=build_lab/$name LIB=../lab/$name.so EXTRA="$flags" &
done
wait
ls -la ../lab
