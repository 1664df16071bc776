#!/bin/bash
# One GPU-box pass: parity tests, bench line, OSNet / YOLO forward timings (profiles/r02_*).
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt )
tail -4 gpurun_out/r02_pytest_gpu.txt
timeout 300 python scripts/time_osnet.py 200 1.0 2>&1 | tail -1 | tee gpurun_out/r02_time_osnet.txt
FM_CONV_TMA=0 timeout 300 python scripts/time_osnet.py 200 1.0 2>&1 | tail -1 | tee -a gpurun_out/r02_time_osnet.txt
timeout 200 python scripts/time_yolo.py 2>&1 | tail -1 | tee gpurun_out/r02_time_yolo.txt
timeout 600 python bench.py --steps 40 --warmup 10 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_c3.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'detector_frame_ms')}, d['e2e']['value'], d['tracker_only'], d['roofline']['ms_per_launch'], d['roofline_tensor']['ms_per_launch'])
PY
