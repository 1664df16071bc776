#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt )
tail -12 gpurun_out/r02_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c3_quick.json 2> gpurun_out/r02_bench_c3_quick.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_c3_quick.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_c3_quick.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'detector_frame_ms', 'gpu_launches')}, d['e2e']['value'], d['tracker_only'], d['roofline']['ms_per_launch'], d['roofline_tensor']['ms_per_launch'])
print(d['repeats'])
for s in d['roofline_stages']: print(s['stage'], s['ms_per_call'], s['calls'], s['ms_per_step'])
PY
