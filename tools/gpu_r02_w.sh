#!/bin/bash
# bodies per region chosen by capacity: the four ensemble sizes + the GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for nb in 4096 3000 2048 1024 512 300; do
( NB=$nb timeout 300 python tools/quick_time.py config4 32 0 4 > gpurun_out/r02_w_time_config4_$nb.txt 2>&1 )
echo "config4 $nb bodies: $(tail -1 gpurun_out/r02_w_time_config4_$nb.txt | cut -c1-60) $(head -1 gpurun_out/r02_w_time_config4_$nb.txt | grep -o "'n_regions': [0-9]*")"
done
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_w_pytest_gpu.log 2>&1 ); tail -3 gpurun_out/r02_w_pytest_gpu.log
