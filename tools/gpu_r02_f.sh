#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "exit $?" >> gpurun_out/r02_bench.err ); tail -3 gpurun_out/r02_bench.err
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err ); tail -2 gpurun_out/r02_bench_reference_arm.err
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; echo "exit $?" >> gpurun_out/r02_smoke.txt ); tail -3 gpurun_out/r02_smoke.txt
