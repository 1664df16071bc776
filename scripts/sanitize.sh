#!/bin/bash
# compute-sanitizer passes over the -m gpu tests (run on the GPU box; summaries land in gpurun_out/).
# usage: scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck] [pytest args...]
tool=${1:-memcheck}; shift
out=gpurun_out/r2_sanitizer_${tool}.txt
sel=${@:-tests/test_gpu_primitives.py tests/test_gpu_osnet_fused.py tests/test_gpu_assoc.py tests/test_gpu_detect.py tests/test_gpu_klt.py}
timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
    python -m pytest $sel -q -x -p no:cacheprovider > $out 2>&1
echo "== $tool: $(grep -c 'Invalid\|Race\|hazard\|Barrier error\|Uninitialized' $out) flagged lines; tail:" 
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $out | tail -5
