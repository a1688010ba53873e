#!/bin/bash
# One gpurun call: micro-benchmarks, GPU parity tests, the bench line, the ncu launch list, step traces.
# Every stage has its own timeout and log so that a late stage cannot hide an early one.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
date +%s > gpurun_out/t0.txt
( timeout 120 tools/bench_math.bin > gpurun_out/bench_math.txt 2>&1; echo "exit $?" >> gpurun_out/bench_math.txt ) 
( timeout 600 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
date +%s > gpurun_out/t1.txt
( timeout 240 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?" >> gpurun_out/bench.err )
( timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "exit $?" >> gpurun_out/smoke.txt )
( timeout 180 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1 )
( timeout 90 python tools/trace_steps.py config3 > gpurun_out/trace_config3.txt 2>&1 )
( NB=4096 timeout 120 python tools/quick_time.py config4 32 2 3 > gpurun_out/quick_config4_persistent.txt 2>&1 )
date +%s > gpurun_out/t2.txt
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | cut -c1-400; tail -5 gpurun_out/bench_math.txt
