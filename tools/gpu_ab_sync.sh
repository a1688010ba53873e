#!/bin/bash
# A/B of the two step-synchronisation changes of the resident kernel on the bench workload, then the
# GPU tests and the bench line with both on (the new default).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rot in 0 1; do for pf in 0 1; do
  ( SBSB200_ROTATE_ITEMS=$rot SBSB200_POLL_FIRST=$pf timeout 120 python tools/quick_time.py config3 32 0 6 > gpurun_out/ab_config3_rot${rot}_pollfirst${pf}.txt 2>&1 )
done; done
export SBSB200_ROTATE_ITEMS=1 SBSB200_POLL_FIRST=1
( timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_b.log )
( timeout 240 python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "exit $?" >> gpurun_out/bench_b.err )
( timeout 90 python tools/trace_steps.py config3 > gpurun_out/trace_config3_b.txt 2>&1 )
unset SBSB200_ROTATE_ITEMS SBSB200_POLL_FIRST
for rot in 0 1; do
  ( SBSB200_ROTATE_ITEMS=$rot SBSB200_POLL_FIRST=$rot timeout 200 python tools/quick_time.py config5 32 0 4 > gpurun_out/ab_config5_new${rot}.txt 2>&1 )
  ( SBSB200_ROTATE_ITEMS=$rot SBSB200_POLL_FIRST=$rot timeout 100 python tools/quick_time.py config2 32 0 6 > gpurun_out/ab_config2_new${rot}.txt 2>&1 )
done
grep -h "frame [3-5]" gpurun_out/ab_config3_*.txt; tail -2 gpurun_out/pytest_gpu_b.log; cut -c1-300 gpurun_out/bench_b.json
