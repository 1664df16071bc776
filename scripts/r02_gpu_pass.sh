#!/bin/bash
# One GPU-box pass: parity tests, bench lines, ncu launch list, OSNet DRAM traffic capture.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt )
tail -3 gpurun_out/r02_pytest_gpu.txt
timeout 600 python bench.py --steps 40 --warmup 10 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02_bench_c3.json
timeout 300 python scripts/time_osnet.py 200 1.0 2>&1 | tail -1 | tee gpurun_out/r02_time_osnet.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 10 --warmup 5 --repeats 1 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
python scripts/agg_launches.py gpurun_out/r02_launches_bench.csv 50 > gpurun_out/r02_launch_summary.txt; tail -3 gpurun_out/r02_launch_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_osnet_dram.csv python scripts/time_osnet.py 200 1.0 --eager > gpurun_out/r02_osnet_dram.log 2>&1
python scripts/osnet_traffic.py gpurun_out/r02_osnet_dram.csv 4 gpurun_out/r02_osnet_traffic
