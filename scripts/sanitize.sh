#!/bin/bash
# compute-sanitizer passes over -m gpu tests (run on the GPU box; logs land in gpurun_out/).
# usage: scripts/sanitize.sh <memcheck|racecheck|synccheck|initcheck> <timeout_s> [pytest args...]
tool=${1:-memcheck}; tmo=${2:-600}; shift; shift
out=gpurun_out/r02_sanitizer_${tool}.txt
sel="${*:-tests/test_gpu_primitives.py tests/test_gpu_assoc.py tests/test_gpu_detect.py}"
eval "timeout $tmo compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 0 python -m pytest $sel -q -p no:cacheprovider" > $out 2>&1
echo "== $tool rc=$? ($sel)"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $out | tail -4
grep -E "Invalid|Race reported|hazard|Barrier error|Uninitialized|Error:" $out | sort | uniq -c | sort -rn | head -8
