#!/bin/bash
# compute-sanitizer passes over the kernels added in round 2 (small sizes): memcheck on a broad subset, racecheck on the kernels whose
# synchronisation is barrier-only (secular / GEMM / back-transformation / stiffness).  Output -> gpurun_out/r02_sanitize.log
out=gpurun_out/r02_sanitize.log
: > $out
run() { echo "=== $*" >> $out; timeout 900 "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Hazard|Invalid|error" | head -20 >> $out; }
run compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fast_update_chain or secular or stiffness_gpu or histories or fsector or (eigenvectors_and_ipr and 8-4) or (step_graph and cubic2d-8) or chain_1d"
run compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(fast_update_chain and cubic2d-8-4.0-4.0-0.5) or (secular_update_stage and 64) or (stiffness_gpu and cubic2d-8) or (eigenvectors_and_ipr and 8-4)"
run compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(fast_update_chain and cubic2d-8-4.0-4.0-0.5) or (two_stage and 256) or (eigenvectors_and_ipr and 8-4)"
cat $out
