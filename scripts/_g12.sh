mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_assoc.py tests/test_gpu_tracker_seq.py tests/test_gpu_mot.py tests/test_gpu_bench_shapes.py tests/test_gpu_klt.py -x -q -p no:cacheprovider > gpurun_out/r02_cascade_tests.txt 2>&1; tail -8 gpurun_out/r02_cascade_tests.txt
timeout 300 python scripts/time_osnet.py 200 1.0 2>&1 | tail -1
timeout 600 python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/r02_bench_c3b.json 2> gpurun_out/r02_bench_c3b.err; echo "bench rc=$?"
FM_FUSE_CASCADE=0 timeout 600 python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/r02_bench_c3_nofuse.json 2> /dev/null
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_c3b.json', 'gpurun_out/r02_bench_c3_nofuse.json'):
    d = json.load(open(f))
    print(f, {k: d[k] for k in ('value', 'ms_per_step', 'detector_frame_ms', 'gpu_launches')}, d['e2e']['value'], d['tracker_only']['ms_per_frame'], d['roofline']['ms_per_launch'], d['roofline_tensor']['ms_per_launch'])
    print([(s['stage'], s['ms_per_call'], s['calls']) for s in d['roofline_stages'] if s['stage'] in ('cost', 'lsa', 'cost+lsa', 'kalman', 'decode+nms')])
PY
