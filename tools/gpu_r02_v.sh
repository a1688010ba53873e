#!/bin/bash
# ensembles: bodies per region 1..5 forced (variant builds) against the heuristic, for 512 / 1024 / 2048 / 4096 bodies; one call
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for nb in 4096 2048 1024 512; do
for g in auto 1 2 3 4 5; do
lib=""; [ "$g" != "auto" ] && lib=$PWD/tools/variants/libsbsb200_bpr$g.so
( SBSB200_LIB=$lib NB=$nb timeout 300 python tools/quick_time.py config4 32 0 4 > gpurun_out/r02_v_time_config4_${nb}_g$g.txt 2>&1 )
echo "config4 $nb bodies, $g per region: $(tail -1 gpurun_out/r02_v_time_config4_${nb}_g$g.txt | cut -c1-60) $(head -1 gpurun_out/r02_v_time_config4_${nb}_g$g.txt | grep -o "'n_regions': [0-9]*")"
done; done
