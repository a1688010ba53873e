#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ce in 0 1; do
  ( SBSB200_CANONICAL_ENTRIES=$ce timeout 120 python tools/quick_time.py config3 32 0 6 > gpurun_out/ab_config3_canonical${ce}.txt 2>&1 )
done
( timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_c.log )
( timeout 240 python bench.py > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "exit $?" >> gpurun_out/bench_c.err )
( timeout 90 python tools/trace_steps.py config3 > gpurun_out/trace_config3_c.txt 2>&1 )
for ce in 0 1; do
  ( SBSB200_CANONICAL_ENTRIES=$ce timeout 100 python tools/quick_time.py config2 32 0 5 > gpurun_out/ab_config2_canonical${ce}.txt 2>&1 )
  ( SBSB200_CANONICAL_ENTRIES=$ce timeout 200 python tools/quick_time.py config5 32 0 4 > gpurun_out/ab_config5_canonical${ce}.txt 2>&1 )
done
grep -h "frame [3-5]" gpurun_out/ab_config3_canonical*.txt; tail -2 gpurun_out/pytest_gpu_c.log; cut -c1-300 gpurun_out/bench_c.json; grep -h "frame 3" gpurun_out/ab_config[25]_canonical*.txt
