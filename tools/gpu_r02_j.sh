#!/bin/bash
# private warps that share a sub-partition with an exchanging warp start late: 0 / 500 / 900 / 1300 ns
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python tools/quick_time.py config3 32 0 6 > gpurun_out/r02_j_time_config3_delay0.txt 2>&1 ); echo "config3 delay0: $(tail -1 gpurun_out/r02_j_time_config3_delay0.txt)"
for d in 500 900 1300; do
( SBSB200_LIB=$PWD/tools/variants/libsbsb200_delay$d.so timeout 300 python tools/quick_time.py config3 32 0 6 > gpurun_out/r02_j_time_config3_delay$d.txt 2>&1 ); echo "config3 delay$d: $(tail -1 gpurun_out/r02_j_time_config3_delay$d.txt)"
done
( SBSB200_LIB=$PWD/tools/variants/libsbsb200_delay900.so REGION_SHAPE=1 timeout 120 python tools/trace_steps.py config3 > gpurun_out/r02_j_trace_config3_delay900.txt 2>&1 )
( SBSB200_LIB=$PWD/tools/variants/libsbsb200_delay900.so timeout 300 python tools/quick_time.py config2 32 0 6 > gpurun_out/r02_j_time_config2_delay900.txt 2>&1 ); echo "config2 delay900: $(tail -1 gpurun_out/r02_j_time_config2_delay900.txt)"
