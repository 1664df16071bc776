#!/bin/bash
mkdir -p gpurun_out
timeout 330 python bench.py --config 4 --steps 20 --warmup 5 --cpu-steps 8 > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; echo "bench c4 rc=$?"; tail -3 gpurun_out/r02_bench_c4.err
