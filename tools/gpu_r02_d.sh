#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log )
tail -3 gpurun_out/r02_pytest_gpu.log
bash tools/gpu_r02_e.sh
( NB=512 timeout 600 python tools/quick_time.py config4 32 0 4 > gpurun_out/r02_time_config4_512.txt 2>&1 ); tail -1 gpurun_out/r02_time_config4_512.txt
