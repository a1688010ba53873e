#!/bin/bash
# early polls (asked for before the last tet of the current cluster) against polls after the pushes: one call, one box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { ( SBSB200_LIB=$2 timeout 300 python tools/quick_time.py $3 $4 0 6 > gpurun_out/r02_u_time_$3_$4_$1.txt 2>&1 ); echo "$3 fp$4 $1: $(tail -1 gpurun_out/r02_u_time_$3_$4_$1.txt)"; }
for rep in a b; do
run early_$rep "" config3 32
run late_$rep $PWD/tools/variants/libsbsb200_noearly.so config3 32
done
run early "" config2 32
run late $PWD/tools/variants/libsbsb200_noearly.so config2 32
run early "" config5 32
run late $PWD/tools/variants/libsbsb200_noearly.so config5 32
( REGION_SHAPE=1 timeout 120 python tools/trace_steps.py config3 > gpurun_out/r02_u_trace_config3.txt 2>&1 )
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_u_pytest_gpu.log 2>&1 ); tail -3 gpurun_out/r02_u_pytest_gpu.log
