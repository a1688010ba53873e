#!/bin/bash
# round 2, first GPU call: contact scenes, region-shape A/B, fast-math A/B + GPU suite on the fast-math build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 400 python tools/explore_r02.py contacts > gpurun_out/r02_contacts.txt 2>&1; echo "exit $?" >> gpurun_out/r02_contacts.txt )
( timeout 400 python tools/explore_r02.py ab > gpurun_out/r02_ab_regions.txt 2>&1; echo "exit $?" >> gpurun_out/r02_ab_regions.txt )
( SBSB200_HANDOFF=1 SBSB200_SLABS=1 timeout 120 python tools/trace_steps.py config3 > gpurun_out/r02_trace_handoff_slabs.txt 2>&1 )
cp soft-body-simulator_b200/lib/libsbsb200.so /tmp/libsbsb200_default.so
cp soft-body-simulator_b200/lib/libsbsb200_fast.so soft-body-simulator_b200/lib/libsbsb200.so
( timeout 200 python tools/quick_time.py config3 32 0 6 > gpurun_out/r02_fastmath_config3.txt 2>&1 )
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_fastmath.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_fastmath.log )
cp /tmp/libsbsb200_default.so soft-body-simulator_b200/lib/libsbsb200.so
( timeout 200 python tools/quick_time.py config3 32 0 6 > gpurun_out/r02_default_config3.txt 2>&1 )
cat gpurun_out/r02_contacts.txt gpurun_out/r02_ab_regions.txt; tail -3 gpurun_out/r02_pytest_gpu_fastmath.log; tail -4 gpurun_out/r02_fastmath_config3.txt gpurun_out/r02_default_config3.txt
