#!/bin/bash
# ensembles: six and seven bodies per region (one CTA of 320 / 352 threads per SM, the 168-register instantiation)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for nb in 4096 2048; do
for g in auto 6 7; do
lib=""; [ "$g" != "auto" ] && lib=$PWD/tools/variants/libsbsb200_bpr$g.so
( SBSB200_LIB=$lib NB=$nb timeout 300 python tools/quick_time.py config4 32 0 4 > gpurun_out/r02_x_time_config4_${nb}_g$g.txt 2>&1 )
echo "config4 $nb bodies, $g per region: $(tail -1 gpurun_out/r02_x_time_config4_${nb}_g$g.txt | cut -c1-60) $(head -1 gpurun_out/r02_x_time_config4_${nb}_g$g.txt | grep -o "'n_regions': [0-9]*")"
done; done
