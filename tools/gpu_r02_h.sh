#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for shape in 1 2; do
( REGION_SHAPE=$shape timeout 300 python tools/quick_time.py config3 32 0 5 > gpurun_out/r02_time_config3_shape$shape.txt 2>&1 ); echo "shape $shape: $(tail -1 gpurun_out/r02_time_config3_shape$shape.txt)"; head -1 gpurun_out/r02_time_config3_shape$shape.txt | grep -o "'n_regions': [0-9]*"
done
