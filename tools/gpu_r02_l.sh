#!/bin/bash
# config 5 / config 3 on whatever box this is, with the clocks and the SM / die layout next to it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,serial,uuid,clocks.sm,clocks.mem,clocks.max.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv > gpurun_out/r02_l_smi.txt 2>&1
( timeout 300 python tools/quick_time.py config5 32 0 5 > gpurun_out/r02_l_time_config5.txt 2>&1 ); echo "config5: $(tail -1 gpurun_out/r02_l_time_config5.txt)"
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv >> gpurun_out/r02_l_smi.txt 2>&1
( timeout 300 python tools/quick_time.py config3 32 0 6 > gpurun_out/r02_l_time_config3.txt 2>&1 ); echo "config3: $(tail -1 gpurun_out/r02_l_time_config3.txt)"
( timeout 60 ./tools/bench_die.bin > gpurun_out/r02_l_bench_die.txt 2>&1 ); grep "SM sides" gpurun_out/r02_l_bench_die.txt
cat gpurun_out/r02_l_smi.txt
