#!/bin/bash
# colour classes of equal size for bodies that live in one region (Kempe chains): ensembles and small scenes, one call = one box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { ( SBSB200_LIB=$2 NB=$5 timeout 300 python tools/quick_time.py $3 $4 0 5 > gpurun_out/r02_r_time_$3$5_$4_$1.txt 2>&1 ); echo "$3 $5 fp$4 $1: $(tail -1 gpurun_out/r02_r_time_$3$5_$4_$1.txt)"; }
for rep in a b; do
run balanced_$rep "" config4 32 4096
run firstfit_$rep $PWD/tools/variants/libsbsb200_nobalance.so config4 32 4096
done
run balanced "" config4 32 512
run firstfit $PWD/tools/variants/libsbsb200_nobalance.so config4 32 512
run balanced "" config4 32 1024
run firstfit $PWD/tools/variants/libsbsb200_nobalance.so config4 32 1024
run balanced "" config1 32
run firstfit $PWD/tools/variants/libsbsb200_nobalance.so config1 32
run balanced "" config4 64 1024
run firstfit $PWD/tools/variants/libsbsb200_nobalance.so config4 64 1024
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_r_pytest_gpu.log 2>&1 ); tail -3 gpurun_out/r02_r_pytest_gpu.log
