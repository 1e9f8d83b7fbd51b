#!/bin/bash
# A/B of library builds (gpurun_variants/lib_*.so) on the two ray workloads
for v in "$@"; do
  echo "== $v"
  PB2_LIB_PATH=$PWD/gpurun_variants/lib_$v.so python harness/prof.py rays_terrain 5 2>&1 | grep "rays_terrain"
  PB2_LIB_PATH=$PWD/gpurun_variants/lib_$v.so python harness/prof.py rays_sphere1m 5 2>&1 | grep "rays_sphere"
done
