#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( REGION_SHAPE=1 timeout 300 python tools/quick_time.py config3 32 0 5 > gpurun_out/r02_time_config3.txt 2>&1 ); tail -1 gpurun_out/r02_time_config3.txt
( timeout 300 python tools/quick_time.py config2 32 0 5 > gpurun_out/r02_time_config2.txt 2>&1 ); tail -1 gpurun_out/r02_time_config2.txt
( timeout 300 python tools/quick_time.py config1 32 0 5 > gpurun_out/r02_time_config1.txt 2>&1 ); tail -1 gpurun_out/r02_time_config1.txt
( NB=4096 timeout 600 python tools/quick_time.py config4 32 0 4 > gpurun_out/r02_time_config4.txt 2>&1 ); tail -1 gpurun_out/r02_time_config4.txt
( timeout 300 python tools/quick_time.py config5 32 0 4 > gpurun_out/r02_time_config5.txt 2>&1 ); tail -1 gpurun_out/r02_time_config5.txt
( timeout 300 python tools/quick_time.py config3 64 0 3 > gpurun_out/r02_time_config3_fp64.txt 2>&1 ); tail -1 gpurun_out/r02_time_config3_fp64.txt
