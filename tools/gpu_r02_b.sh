#!/bin/bash
# round 2: first run of the resident kernel (guest slots, pencil regions): GPU suite, timings of every config, step trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log )
tail -5 gpurun_out/r02_pytest_gpu.log
for cfg in config3 config2 config1 config5; do
  ( timeout 300 python tools/quick_time.py $cfg 32 0 5 > gpurun_out/r02_time_$cfg.txt 2>&1 ); tail -3 gpurun_out/r02_time_$cfg.txt
done
( NB=4096 timeout 600 python tools/quick_time.py config4 32 0 4 > gpurun_out/r02_time_config4.txt 2>&1 ); tail -3 gpurun_out/r02_time_config4.txt
( REGION_SHAPE=1 timeout 300 python tools/quick_time.py config3 32 0 5 > gpurun_out/r02_time_config3_compact.txt 2>&1 ); tail -2 gpurun_out/r02_time_config3_compact.txt
( timeout 120 python tools/trace_steps.py config3 > gpurun_out/r02_trace_config3.txt 2>&1 ); tail -16 gpurun_out/r02_trace_config3.txt
( REGION_SHAPE=1 timeout 120 python tools/trace_steps.py config3 > gpurun_out/r02_trace_config3_compact.txt 2>&1 ); tail -16 gpurun_out/r02_trace_config3_compact.txt
