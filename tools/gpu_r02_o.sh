#!/bin/bash
# in-place constraint removal: parity tests, cost at the bench size; whole GPU suite; config 5 (and an ncu capture when this box is a slow one)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_cpp_facade.py -m gpu -x -q -k "removed" > gpurun_out/r02_o_pytest_removed.log 2>&1 ); tail -5 gpurun_out/r02_o_pytest_removed.log
( timeout 300 python tools/time_remove.py > gpurun_out/r02_time_remove_constraints.txt 2>&1 ); cat gpurun_out/r02_time_remove_constraints.txt | tail -5
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_o_pytest_gpu.log 2>&1 ); tail -3 gpurun_out/r02_o_pytest_gpu.log
nvidia-smi --query-gpu=serial,memory.used,ecc.mode.current,clocks.mem --format=csv > gpurun_out/r02_o_smi.txt; nvidia-smi >> gpurun_out/r02_o_smi.txt
( timeout 300 python tools/quick_time.py config5 32 0 4 > gpurun_out/r02_o_time_config5.txt 2>&1 ); echo "config5: $(tail -1 gpurun_out/r02_o_time_config5.txt)"
ms=$(tail -1 gpurun_out/r02_o_time_config5.txt | awk '{print int($3)}')
if [ "$ms" -gt 27 ]; then
echo "slow box: ncu capture of config 5"
( timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_substep_resident -s 25 -c 1 -f -o gpurun_out/r02_o_full_config5_slow python tools/quick_time.py config5 32 0 4 > gpurun_out/r02_o_ncu_config5.log 2>&1 ); tail -2 gpurun_out/r02_o_ncu_config5.log
fi
