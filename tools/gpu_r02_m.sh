#!/bin/bash
# one call, one box: slots in vertex order vs spread over the banks; poll intensity (generations in flight, gap, sleep, sentinel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 60 ./tools/bench_die.bin > gpurun_out/r02_m_bench_die.txt 2>&1 ); grep "SM sides" gpurun_out/r02_m_bench_die.txt
run() { ( SBSB200_LIB=$2 timeout 300 python tools/quick_time.py $3 32 0 6 > gpurun_out/r02_m_time_$3_$1.txt 2>&1 ); echo "$3 $1: $(tail -1 gpurun_out/r02_m_time_$3_$1.txt)"; }
for rep in a b; do
run default_$rep "" config3
for v in nospread gen1 gen1sleep gen3gap40 sentinel sentinel1; do run ${v}_$rep $PWD/tools/variants/libsbsb200_$v.so config3; done
done
run default "" config5
run nospread $PWD/tools/variants/libsbsb200_nospread.so config5
run sentinel $PWD/tools/variants/libsbsb200_sentinel.so config5
run default "" config2
run nospread $PWD/tools/variants/libsbsb200_nospread.so config2
run sentinel $PWD/tools/variants/libsbsb200_sentinel.so config2
