#!/bin/bash
# Round-2 final GPU pass (one gpurun call): parity tests, bench lines for BASELINE configs 1-4 + the reference arm,
# ncu launch list of the bench command, one `--set full` capture of the dominant ReID kernel, OSNet DRAM traffic.
# Outputs land in gpurun_out/ (copied to profiles/ afterwards).
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r02_pytest_gpu_final.txt 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu_final.txt )
tail -4 $O/r02_pytest_gpu_final.txt
grep -q "pytest rc=0" $O/r02_pytest_gpu_final.txt || { echo "parity tests failed: stopping the pass"; grep -E "^(FAILED|ERROR)|Error" $O/r02_pytest_gpu_final.txt | head -20; exit 1; }
timeout 900 python bench.py --steps 40 --warmup 10 > $O/r02_bench_c3.json 2> $O/r02_bench_c3.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_c3_driver_args.json 2> /dev/null; echo "bench c3 (20/5) rc=$?"
for c in 1 2 4; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --cpu-steps 10 > $O/r02_bench_c$c.json 2> $O/r02_bench_c$c.err; echo "bench c$c rc=$?"
done
timeout 600 python bench.py --impl reference --no-nets --steps 10 --warmup 2 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 300 python scripts/time_osnet.py 200 1.0 2>&1 | tail -1 | tee $O/r02_time_osnet_final.txt
timeout 200 python scripts/time_yolo.py 2>&1 | tail -1 | tee $O/r02_time_yolo_final.txt
timeout 200 python scripts/time_lsa.py 2>&1 | tail -6 | tee $O/r02_time_lsa_final.txt
# launch list of the bench command (never a bench value: numbers under ncu are serialised and cold)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02_launches_bench_ncu_final.csv \
    python bench.py --steps 10 --warmup 5 --repeats 1 --no-cpu-baseline > $O/r02_bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?"
python scripts/agg_launches.py $O/r02_launches_bench_ncu_final.csv 60 > $O/r02_launch_summary_final.txt 2>&1; head -12 $O/r02_launch_summary_final.txt
# OSNet forward DRAM traffic (eager replay, 2 warm + 2 timed forwards = 4 forwards in the capture)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $O/r02_osnet_dram_final.csv python scripts/time_osnet.py 200 1.0 --eager > /dev/null 2>&1; echo "ncu osnet dram rc=$?"
python scripts/osnet_traffic.py $O/r02_osnet_dram_final.csv 4 $O/r02_osnet_final 2>&1 | tail -3
# one full-set capture of the dominant kernel (stage-1 OSBlock streams kernel)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:osb_streams_kernel -c 1 -o $O/r02_osb_streams_full \
    python scripts/time_osnet.py 200 1.0 --eager > /dev/null 2>&1; echo "ncu full rc=$?"
ncu -i $O/r02_osb_streams_full.ncu-rep --page raw --csv > $O/r02_osb_streams_full_raw.csv 2>/dev/null
ls -la $O | tail -30
