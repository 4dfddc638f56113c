#!/bin/bash
# development (B200 box): full ncu captures of convolutions as they run inside a chain (planes in, dense + planes out)
mkdir -p gpurun_out
TAG=${TAG:-chain}
for c in ${CONVS:-ru1x1b ru3x3 cc2}; do
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -s 2 -c 1 -o gpurun_out/r02_conv_${TAG}_$c -f python tools/prof_conv.py $c planes > gpurun_out/r02_ncu_conv_${TAG}_$c.log 2>&1
  ncu -i gpurun_out/r02_conv_${TAG}_$c.ncu-rep --page raw --csv > gpurun_out/r02_conv_${TAG}_${c}_ncu_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_conv_${TAG}_$c.ncu-rep --page source --csv > gpurun_out/r02_conv_${TAG}_${c}_ncu_src.csv 2>/dev/null
  rm -f gpurun_out/r02_conv_${TAG}_$c.ncu-rep
done
ls -la gpurun_out/r02_conv_${TAG}_* | cut -c1-150
