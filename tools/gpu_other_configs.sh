#!/bin/bash
# bench lines of the other single-GPU workloads on the final build of the round
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in config1 config2 config5 config4; do
  ( timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_final_$w.json 2> gpurun_out/bench_final_$w.err )
  cut -c1-230 gpurun_out/bench_final_$w.json
done
