#!/bin/bash
# hunt for a slow box: time config 5 on one GPU; when it is slow, capture what differs (ncu of its substep kernel, nvidia-smi -q,
# the die micro-benchmark, copy bandwidth and DRAM latency)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$(date +%H%M%S)
( timeout 300 python tools/quick_time.py config5 32 0 3 > gpurun_out/r02_s_time_config5_$tag.txt 2>&1 ); echo "config5: $(tail -1 gpurun_out/r02_s_time_config5_$tag.txt)"
ms=$(tail -1 gpurun_out/r02_s_time_config5_$tag.txt | awk '{print int($3)}')
nvidia-smi --query-gpu=serial,uuid,pci.bus_id,memory.used,ecc.mode.current,clocks.mem,clocks.sm --format=csv | tee gpurun_out/r02_s_smi_$tag.txt
if [ "$ms" -gt 27 ]; then
echo "SLOW BOX"
nvidia-smi -q > gpurun_out/r02_s_smi_q_slow_$tag.txt 2>&1
( timeout 60 ./tools/bench_die.bin > gpurun_out/r02_s_bench_die_slow_$tag.txt 2>&1 )
( timeout 120 ./tools/bench_lat.bin > gpurun_out/r02_s_bench_lat_slow_$tag.txt 2>&1 )
( timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_substep_resident -s 15 -c 1 -f -o gpurun_out/r02_s_full_config5_slow python tools/quick_time.py config5 32 0 3 > gpurun_out/r02_s_ncu_config5_slow.log 2>&1 ); tail -2 gpurun_out/r02_s_ncu_config5_slow.log
( timeout 300 python tools/quick_time.py config3 32 0 5 > gpurun_out/r02_s_time_config3_slow_$tag.txt 2>&1 ); echo "config3: $(tail -1 gpurun_out/r02_s_time_config3_slow_$tag.txt)"
fi
