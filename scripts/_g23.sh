#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r02_pytest_gpu_final2.txt 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu_final2.txt )
tail -4 $O/r02_pytest_gpu_final2.txt
grep -q "pytest rc=0" $O/r02_pytest_gpu_final2.txt || { echo "parity tests failed"; grep -E "^(FAILED|ERROR)|Error|assert" $O/r02_pytest_gpu_final2.txt | head -20; exit 1; }
timeout 300 python bench.py --config 1 --steps 20 --warmup 5 --cpu-steps 10 > $O/r02_bench_c1.json 2> $O/r02_bench_c1.err; echo "bench c1 rc=$?"; tail -2 $O/r02_bench_c1.err
timeout 400 python bench.py --steps 40 --warmup 10 > $O/r02_bench_c3_v2.json 2> $O/r02_bench_c3_v2.err; echo "bench c3 rc=$?"
timeout 500 python bench.py --config 4 --steps 20 --warmup 5 --cpu-steps 10 > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err; echo "bench c4 rc=$?"; tail -2 $O/r02_bench_c4.err
