#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in libsbsb200.so libsbsb200_sleep50.so libsbsb200_sleep200.so libsbsb200.so; do
    ( SBSB200_LIB=$PWD/soft-body-simulator_b200/lib/$lib timeout 600 python tools/quick_time.py config3 32 0 6 > gpurun_out/r02_ab_sleep_${lib}.txt 2>&1 ); echo "$lib: $(tail -1 gpurun_out/r02_ab_sleep_${lib}.txt)"
done
