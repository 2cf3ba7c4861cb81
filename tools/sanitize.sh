#!/bin/bash
# Race / memory checks of the hand-written kernels (the reference has none — SURVEY §5.2).
# Run on a GPU box:   gpurun --timeout 900 -- bash tools/sanitize.sh
# memcheck + racecheck on the single-rank kernels; the multi-rank worlds simulated on
# one GPU spin on flags and are excluded (the sanitizer serialises kernels).
set -x
export PYTHONPATH=$(dirname "$0")/..
SAN=/usr/local/cuda/bin/compute-sanitizer
$SAN --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fused.py -x -q -k "not 8192" 2>&1 | tail -5
$SAN --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_gemm.py -x -q -k "128-128-64" 2>&1 | tail -5
$SAN --tool memcheck --error-exitcode 1 python -m pytest "tests/test_gpu_sparse.py::test_push_claim_apply" -x -q -k "1-HYBRID" 2>&1 | tail -5
