#!/bin/bash
# development (B200 box): per-layer convolution timings + full ncu captures of representative convolutions
mkdir -p gpurun_out
python tools/conv_time.py > gpurun_out/r02_conv_time.log 2>&1 || { tail -5 gpurun_out/r02_conv_time.log; exit 1; }
cat gpurun_out/r02_conv_time.log
for c in ${CONVS:-ru1x1b ru3x3 dse3x3 cc2}; do
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -s 2 -c 1 -o gpurun_out/r02_conv_$c -f python tools/prof_conv.py $c > gpurun_out/r02_ncu_conv_$c.log 2>&1
  ncu -i gpurun_out/r02_conv_$c.ncu-rep --page raw --csv > gpurun_out/r02_conv_${c}_ncu_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_conv_$c.ncu-rep --page source --csv > gpurun_out/r02_conv_${c}_ncu_src.csv 2>/dev/null
  rm -f gpurun_out/r02_conv_$c.ncu-rep
done
ls -la gpurun_out/r02_conv_* | cut -c1-150
