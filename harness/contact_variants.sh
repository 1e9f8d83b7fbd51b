#!/bin/bash
# A/B of library builds (gpurun_variants/lib_*.so, built on the dev box) on the 2^22 hull-pair contact workload.
for v in "$@"; do
  echo "== $v"
  PB2_LIB_PATH=$PWD/gpurun_variants/lib_$v.so python harness/prof.py contacts 3 2>&1 | grep -v "^$"
done
