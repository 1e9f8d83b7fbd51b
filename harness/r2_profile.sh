#!/bin/bash
# One gpurun call: ncu evidence for the round-2 kernels (profiles/README.md says how each file is read).
#   gpurun --timeout 1500 -- 'bash harness/r2_profile.sh'
set -u
out=gpurun_out/r2prof
mkdir -p "$out"
# 1. launch list of the default bench (shares only: per-launch times under ncu are cold-cache and serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$out/bench_launches.csv" \
    python bench.py --steps 2 --warmup 3 --skip-cpu > "$out/bench_under_ncu.log" 2>&1
echo "launches rc=$?"
# 2. ray kernel, --set full, terrain 8M triangles / 2^23 rays
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_raycast_wide_shared -s 2 -c 1 -o "$out/rays_terrain" -f \
    python harness/prof.py rays_terrain 2 > "$out/rays_terrain.log" 2>&1
echo "rays rc=$?"
# 3. contact kernels, --set full, 2^22 hull pairs
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_contact_gjk|k_contact_epa2|k_contact_finish' -s 3 -c 3 -o "$out/contacts" -f \
    python harness/prof.py contacts 1 > "$out/contacts.log" 2>&1
echo "contacts rc=$?"
ls -la "$out"
