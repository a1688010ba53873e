#!/bin/bash
# persisting L2 window over the tet slot records of scenes whose per-sweep records exceed the L2 (config 5, config 4): one call
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { ( SBSB200_LIB=$2 NB=4096 timeout 300 python tools/quick_time.py $3 32 0 5 > gpurun_out/r02_z_time_$3_$1.txt 2>&1 ); echo "$3 $1: $(tail -1 gpurun_out/r02_z_time_$3_$1.txt)"; }
for rep in a b; do
run window_$rep "" config5
run plain_$rep $PWD/tools/variants/libsbsb200_nowindow.so config5
done
run window "" config4
run plain $PWD/tools/variants/libsbsb200_nowindow.so config4
run window "" config3
