#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt )
tail -3 gpurun_out/r02_pytest_gpu.txt
timeout 300 python scripts/profile_step_split.py > gpurun_out/r02_profile_split.txt 2>&1
head -3 gpurun_out/r02_profile_split.txt
