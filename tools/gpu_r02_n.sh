#!/bin/bash
# config 5 on one GPU runs at ~21 ms on some boxes and ~35 ms on others: time it, then one full ncu capture of its substep kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,serial,clocks.sm,clocks.mem,power.draw,temperature.gpu,memory.used,memory.total,ecc.mode.current --format=csv > gpurun_out/r02_n_smi.txt 2>&1
( timeout 300 python tools/quick_time.py config5 32 0 4 > gpurun_out/r02_n_time_config5.txt 2>&1 ); echo "config5: $(tail -1 gpurun_out/r02_n_time_config5.txt)"
( timeout 300 python tools/quick_time.py config3 32 0 5 > gpurun_out/r02_n_time_config3.txt 2>&1 ); echo "config3: $(tail -1 gpurun_out/r02_n_time_config3.txt)"
( timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_substep_resident -s 25 -c 1 -f -o gpurun_out/r02_n_full_config5 python tools/quick_time.py config5 32 0 4 > gpurun_out/r02_n_ncu_config5.log 2>&1 ); tail -2 gpurun_out/r02_n_ncu_config5.log
cat gpurun_out/r02_n_smi.txt
