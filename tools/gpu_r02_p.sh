#!/bin/bash
# tet records of a region in shared memory (TetStore) against the records in global memory: one call, one box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { ( SBSB200_LIB=$2 timeout 300 python tools/quick_time.py $3 $4 0 6 > gpurun_out/r02_p_time_$3_$4_$1.txt 2>&1 ); echo "$3 fp$4 $1: $(tail -1 gpurun_out/r02_p_time_$3_$4_$1.txt)"; }
for rep in a b; do
run local_$rep "" config3 32
run global_$rep $PWD/tools/variants/libsbsb200_nolocal.so config3 32
done
run local "" config2 32
run global $PWD/tools/variants/libsbsb200_nolocal.so config2 32
run local "" config3 64
run global $PWD/tools/variants/libsbsb200_nolocal.so config3 64
run local "" config1 32
( REGION_SHAPE=1 timeout 120 python tools/trace_steps.py config3 > gpurun_out/r02_p_trace_config3.txt 2>&1 )
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_p_pytest_gpu.log 2>&1 ); tail -3 gpurun_out/r02_p_pytest_gpu.log
