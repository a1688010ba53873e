#!/bin/bash
# Round-2 evidence on one GPU: GPU tests, smoke, bench line, reference arm, ncu launch list and full capture of the substep
# kernel, step traces, the other workloads, sanitizer runs.  Every stage has its own timeout and log.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_final_pytest_gpu.log )
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.txt 2>&1; echo "exit $?" >> gpurun_out/r02_final_smoke.txt )
( timeout 600 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "exit $?" >> gpurun_out/r02_final_bench.err )
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_reference_arm.json 2> gpurun_out/r02_final_bench_reference_arm.err )
for w in config1 config2; do
( timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline >> gpurun_out/r02_final_bench_other_configs.jsonl 2>> gpurun_out/r02_final_bench_other_configs.err )
done
( timeout 300 python bench.py --region-shape 0 --steps 5 --warmup 3 --no-cpu-baseline --no-sub >> gpurun_out/r02_final_bench_other_configs.jsonl 2>> gpurun_out/r02_final_bench_other_configs.err )
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-sub > gpurun_out/r02_bench_under_ncu.log 2>&1 )
( timeout 400 ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_substep_resident -s 25 -c 1 -f -o gpurun_out/r02_full_k_substep_resident python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --no-sub > gpurun_out/r02_bench_under_ncu_full.log 2>&1 )
( REGION_SHAPE=1 timeout 120 python tools/trace_steps.py config3 > gpurun_out/r02_final_trace_config3_compact.txt 2>&1 )
( REGION_SHAPE=0 timeout 120 python tools/trace_steps.py config3 > gpurun_out/r02_final_trace_config3_pencils.txt 2>&1 )
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -x -k "region_shapes and 32-1 or removed_in_place and 32 or surface_output_with_colours or ensemble_of_independent" > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "exit $?" >> gpurun_out/r02_sanitizer_memcheck.txt )
( timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -x -k "general_route_counter or ensemble_of_independent or region_shapes and 32" > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "exit $?" >> gpurun_out/r02_sanitizer_racecheck.txt )
( timeout 300 python tools/time_remove.py > gpurun_out/r02_time_remove_constraints.txt 2>&1 )
( timeout 60 ./tools/bench_die.bin > gpurun_out/r02_microbench_die_homes.txt 2>&1 )
tail -3 gpurun_out/r02_final_pytest_gpu.log; tail -3 gpurun_out/r02_final_smoke.txt; cut -c1-300 gpurun_out/r02_final_bench.json; tail -n 3 gpurun_out/r02_sanitizer_memcheck.txt; tail -n 3 gpurun_out/r02_sanitizer_racecheck.txt; ls -la gpurun_out/*.ncu-rep
