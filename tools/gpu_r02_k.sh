#!/bin/bash
# mailboxes homed on the die of their reader, regions dealt to the SMs by die: A/B in one call (process-to-process and box-to-box
# variation is several per cent: only runs of one call compare)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rep in 1 2; do for da in 0 1; do
( DIE_AWARE=$da timeout 300 python tools/quick_time.py config3 32 0 6 > gpurun_out/r02_k_time_config3_die${da}_$rep.txt 2>&1 ); echo "config3 die_aware=$da: $(tail -1 gpurun_out/r02_k_time_config3_die${da}_$rep.txt)"
done; done
grep placement gpurun_out/r02_k_time_config3_die1_1.txt
for cfg in config2 config5; do for da in 0 1; do
( DIE_AWARE=$da timeout 300 python tools/quick_time.py $cfg 32 0 5 > gpurun_out/r02_k_time_${cfg}_die$da.txt 2>&1 ); echo "$cfg die_aware=$da: $(tail -1 gpurun_out/r02_k_time_${cfg}_die$da.txt)"
done; done
grep placement gpurun_out/r02_k_time_config5_die1.txt gpurun_out/r02_k_time_config2_die1.txt
( REGION_SHAPE=1 timeout 120 python tools/trace_steps.py config3 > gpurun_out/r02_k_trace_config3.txt 2>&1 )
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_k_pytest_gpu.log 2>&1 ); tail -3 gpurun_out/r02_k_pytest_gpu.log
