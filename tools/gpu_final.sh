#!/bin/bash
# Final evidence of the round: exchange-pattern probe, GPU tests, smoke, bench line, reference arm, ncu launch
# list and one full capture of the substep kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 60 tools/bench_lat.bin > gpurun_out/bench_lat.txt 2>&1; echo "exit $?" >> gpurun_out/bench_lat.txt )
( timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log )
( timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.txt 2>&1; echo "exit $?" >> gpurun_out/smoke_final.txt )
( timeout 240 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "exit $?" >> gpurun_out/bench_final.err )
( timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err )
( timeout 180 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu_final.log 2>&1 )
( timeout 240 ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_substep_persistent -s 25 -c 1 -f -o gpurun_out/full_k_substep_persistent python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu_full.log 2>&1 )
( timeout 90 python tools/trace_steps.py config3 > gpurun_out/trace_config3_final.txt 2>&1 )
tail -2 gpurun_out/pytest_gpu_final.log; cut -c1-330 gpurun_out/bench_final.json; grep exchange gpurun_out/bench_lat.txt; ls -la gpurun_out/*.ncu-rep
