#!/bin/bash
# ncu capture of the island (ensemble) instantiation of the substep kernel on config 4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( NB=4096 timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_substep_resident -s 15 -c 1 -f -o gpurun_out/r02_q_full_config4 python tools/quick_time.py config4 32 0 3 > gpurun_out/r02_q_ncu_config4.log 2>&1 ); tail -3 gpurun_out/r02_q_ncu_config4.log
