#!/bin/bash
# timing experiment (results are wrong on purpose): pulls that do not wait -> what a frame would cost with a free exchange
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python tools/quick_time.py config3 32 0 5 > gpurun_out/r02_t_time_config3_wait.txt 2>&1 ); echo "config3 waiting: $(tail -1 gpurun_out/r02_t_time_config3_wait.txt)"
( SBSB200_LIB=$PWD/tools/variants/libsbsb200_nowait.so timeout 300 python tools/quick_time.py config3 32 0 5 > gpurun_out/r02_t_time_config3_nowait.txt 2>&1 ); echo "config3 not waiting: $(tail -1 gpurun_out/r02_t_time_config3_nowait.txt)"
( SBSB200_LIB=$PWD/tools/variants/libsbsb200_nowait.so timeout 300 python tools/quick_time.py config2 32 0 5 > gpurun_out/r02_t_time_config2_nowait.txt 2>&1 ); echo "config2 not waiting: $(tail -1 gpurun_out/r02_t_time_config2_nowait.txt)"
( timeout 300 python tools/quick_time.py config2 32 0 5 > gpurun_out/r02_t_time_config2_wait.txt 2>&1 ); echo "config2 waiting: $(tail -1 gpurun_out/r02_t_time_config2_wait.txt)"
