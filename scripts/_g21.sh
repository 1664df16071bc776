#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_onnx_import.py tests/test_gpu_detect.py -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_onnx.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_onnx.txt )
tail -30 gpurun_out/r02_pytest_onnx.txt
timeout 300 python scripts/profile_step_split.py > gpurun_out/r02_profile_split.txt 2>&1
head -3 gpurun_out/r02_profile_split.txt
