#!/bin/bash
# slots spread over the shared-memory banks + polls in flight (1 / 2 / 3 generations): timing of the workloads + GPU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in gen1 gen2; do
( SBSB200_LIB=$PWD/tools/variants/libsbsb200_$g.so timeout 300 python tools/quick_time.py config3 32 0 6 > gpurun_out/r02_i_time_config3_$g.txt 2>&1 ); echo "config3 $g: $(tail -1 gpurun_out/r02_i_time_config3_$g.txt)"
done
( SBSB200_LIB=$PWD/tools/variants/libsbsb200_gen1.so timeout 300 python tools/quick_time.py config5 32 0 4 > gpurun_out/r02_i_time_config5_gen1.txt 2>&1 ); echo "config5 gen1: $(tail -1 gpurun_out/r02_i_time_config5_gen1.txt)"
for cfg in config3 config2 config1 config5; do
( timeout 300 python tools/quick_time.py $cfg 32 0 6 > gpurun_out/r02_i_time_$cfg.txt 2>&1 ); echo "$cfg: $(tail -1 gpurun_out/r02_i_time_$cfg.txt)"
done
( NB=512 timeout 300 python tools/quick_time.py config4 32 0 5 > gpurun_out/r02_i_time_config4_512.txt 2>&1 ); echo "config4/512: $(tail -1 gpurun_out/r02_i_time_config4_512.txt)"
( NB=4096 timeout 300 python tools/quick_time.py config4 32 0 4 > gpurun_out/r02_i_time_config4.txt 2>&1 ); echo "config4: $(tail -1 gpurun_out/r02_i_time_config4.txt)"
( timeout 300 python tools/quick_time.py config3 64 0 4 > gpurun_out/r02_i_time_config3_fp64.txt 2>&1 ); echo "config3 fp64: $(tail -1 gpurun_out/r02_i_time_config3_fp64.txt)"
( REGION_SHAPE=1 timeout 120 python tools/trace_steps.py config3 > gpurun_out/r02_i_trace_config3.txt 2>&1 )
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_i_pytest_gpu.log 2>&1 ); tail -3 gpurun_out/r02_i_pytest_gpu.log
