#!/bin/bash
# compute-sanitizer over the paths added in the second session of round 2 (and the exchange path under racecheck)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -x -k "removed_in_place and 32 or surface_output_with_colours or ensemble_of_independent" > gpurun_out/r02_sanitizer_memcheck_session2.txt 2>&1; echo "exit $?" >> gpurun_out/r02_sanitizer_memcheck_session2.txt ); tail -4 gpurun_out/r02_sanitizer_memcheck_session2.txt
( timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x -k "ensemble_of_independent or region_shapes and 32" > gpurun_out/r02_sanitizer_racecheck_session2.txt 2>&1; echo "exit $?" >> gpurun_out/r02_sanitizer_racecheck_session2.txt ); tail -4 gpurun_out/r02_sanitizer_racecheck_session2.txt
( timeout 300 python tools/quick_time.py config3 32 0 6 > gpurun_out/r02_y_time_config3.txt 2>&1 ); echo "config3: $(tail -1 gpurun_out/r02_y_time_config3.txt)"
