#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in libsbsb200.so libsbsb200_mb3.so libsbsb200_mb4.so; do
  for nb in 4096 512; do
    ( SBSB200_LIB=$PWD/soft-body-simulator_b200/lib/$lib NB=$nb timeout 600 python tools/quick_time.py config4 32 0 4 > gpurun_out/r02_ab_island_${lib}_$nb.txt 2>&1 ); echo "$lib $nb: $(tail -1 gpurun_out/r02_ab_island_${lib}_$nb.txt)"
  done
done
