#!/bin/bash
# First hardware run of the paths written after round 1's GPU budget was spent. One gpurun call:
#   gpurun --timeout 900 -- 'bash harness/first_hw_run.sh'
# Writes everything under gpurun_out/first_hw/. Each step runs under its own timeout so that a hang cannot take the call down.
set -u
out=gpurun_out/first_hw
mkdir -p "$out"
# 1. the late test files, most trusted first (no -x: see every result)
timeout 600 python -m pytest tests/test_zx_cpp_mirror.py tests/test_zy_reference_pins_gpu.py tests/test_zz_manifold_update_gpu.py \
    tests/test_zzz_compound_compound_gpu.py -m gpu -q > "$out/pytest.log" 2>&1
echo "pytest rc=$?" >> "$out/pytest.log"
# 2. memcheck of the new kernels on the small golden scene
timeout 300 compute-sanitizer --tool memcheck python harness/sanitize_probe.py persistence > "$out/memcheck.log" 2>&1
echo "memcheck rc=$?" >> "$out/memcheck.log"
# 3. the bench entry that times them (headline only + the extra entries; CPU baseline skipped)
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu > "$out/bench.json" 2> "$out/bench.err"
echo "bench rc=$?" >> "$out/bench.err"
# 4. launch list of the new kernels (per-launch times are cold-cache and serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_manifold_try_update|k_manifold_match|k_cc_candidates|k_cc_reduce|k_contact_manifolds' \
    -c 60 --csv --log-file "$out/launches_persistence.csv" python harness/sanitize_probe.py persistence > "$out/ncu.log" 2>&1
echo "ncu rc=$?" >> "$out/ncu.log"
tail -3 "$out/pytest.log" "$out/memcheck.log" "$out/bench.err"
