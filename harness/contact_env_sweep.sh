#!/bin/bash
# env-knob sweep on the 2^22 hull-pair contact workload: bash harness/contact_env_sweep.sh "A=1" "A=2 B=3" ...
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python harness/prof.py contacts 3 2>&1 | grep -v "^$\|checksum"
done
