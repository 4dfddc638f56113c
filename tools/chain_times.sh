#!/bin/bash
# development (B200 box): CUDA-event time of convolutions as they run inside a chain (planes in, dense + planes out)
for c in ${CONVS:-ru1x1a ru3x3 ru1x1b cc0 cc2 cc4 dse3x3 x2 dx3 dx4}; do
  timeout -s KILL 120 python tools/prof_conv.py $c planes time 2>&1 | grep median
done
