#!/bin/bash
# One gpurun call: the late parity test, then the ncu captures of the final kernels (profiles/README.md says how each file is read).
#   gpurun --timeout 1400 -- 'bash harness/r2_profile_final.sh'
set -u
out=gpurun_out/r2final
mkdir -p "$out"
timeout 300 python -m pytest tests/test_zzzz_compound_trimesh_gpu.py -x -q > "$out/pytest_ct.log" 2>&1
echo "pytest rc=$?"; tail -5 "$out/pytest_ct.log"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_raycast_wide_shared -s 2 -c 1 -o "$out/rays_terrain" -f \
    python harness/prof.py rays_terrain 2 > "$out/rays_terrain.log" 2>&1
echo "rays rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_contact_gjk|k_contact_epa2|k_contact_finish' -s 4 -c 4 -o "$out/contacts" -f \
    python harness/prof.py contacts 1 > "$out/contacts.log" 2>&1
echo "contacts rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file "$out/bench_launches.csv" \
    python bench.py --steps 2 --warmup 3 --skip-cpu > "$out/bench_under_ncu.log" 2>&1
echo "launches rc=$?"
ls -la "$out"
