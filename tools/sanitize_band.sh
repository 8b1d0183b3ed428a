#!/bin/bash
# compute-sanitizer passes over the band kernel (sb2sb.cu) and the local Chebyshev scheme (kpm2d.cu), small sizes.
# Output -> gpurun_out/r02_sanitize_band.log
out=gpurun_out/r02_sanitize_band.log
: > $out
run() { echo "=== $*" >> $out; timeout 1200 "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Hazard|Invalid|error" | head -20 >> $out; }
run compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(band_path_spectra and cubic2d-16) or band_path_selection or band_path_chain or (kpm_local_matches_full and cubic2d-16) or kpm_local_chain or kpm_local_other"
run compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(band_path_spectra and cubic2d-16) or (kpm_local_matches_full and cubic2d-16)"
run compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(band_path_spectra and cubic2d-16) or (kpm_local_matches_full and cubic2d-16)"
cat $out
