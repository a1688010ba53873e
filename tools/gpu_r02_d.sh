#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log )
tail -3 gpurun_out/r02_pytest_gpu.log
for shape in 1 0; do
( REGION_SHAPE=$shape timeout 300 python tools/quick_time.py config3 32 0 5 > gpurun_out/r02_time_config3_shape$shape.txt 2>&1 ); tail -1 gpurun_out/r02_time_config3_shape$shape.txt
done
( timeout 300 python tools/quick_time.py config2 32 0 5 > gpurun_out/r02_time_config2.txt 2>&1 ); tail -1 gpurun_out/r02_time_config2.txt
( timeout 300 python tools/quick_time.py config1 32 0 5 > gpurun_out/r02_time_config1.txt 2>&1 ); tail -1 gpurun_out/r02_time_config1.txt
( NB=4096 timeout 600 python tools/quick_time.py config4 32 0 4 > gpurun_out/r02_time_config4.txt 2>&1 ); tail -1 gpurun_out/r02_time_config4.txt
( timeout 300 python tools/quick_time.py config5 32 0 4 > gpurun_out/r02_time_config5.txt 2>&1 ); tail -1 gpurun_out/r02_time_config5.txt; head -c 700 gpurun_out/r02_time_config5.txt | tail -c 350
( REGION_SHAPE=1 timeout 120 python tools/trace_steps.py config3 > gpurun_out/r02_trace_config3_compact.txt 2>&1 ); tail -100 gpurun_out/r02_trace_config3_compact.txt | grep -v Warning | grep -v "np.nan" | grep -A9 "last tet done\|ready for\|out of the barrier\|barrier-to"
