#!/bin/bash
# round-2 evidence (run on the B200 box): launch list of the bench command, full ncu captures of the top kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --legs none --no-cpu-baseline --no-e2e --no-clocks"
$CMD > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err || { tail -5 gpurun_out/r02_bench_plain.err; exit 1; }
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_launches.csv $CMD > /dev/null 2>&1
cap() {   # name, kernel regex, launches to skip, launches to capture, command...
  local name=$1 k=$2 s=$3 c=$4; shift 4
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $s -c $c -o gpurun_out/$name -f "$@" > gpurun_out/${name}.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_ncu_raw.csv 2>/dev/null
}
cap r02_attn_sp mwa_sp_kernel 0 2 python tools/prof_one.py attn8 auto blob
cap r02_attn_sp_dense mwa_sp_kernel 0 1 python tools/prof_one.py attn8 auto ones
cap r02_gdn_tc gdn_tc 0 6 python tools/prof_gdn_sites.py
cap r02_attn_small mwa_small_kernel 0 1 python tools/prof_one.py attn4 auto ones
cap r02_copy_dropped mwa_copy_dropped 0 1 python tools/prof_one.py attn8 auto blob
for c in ru3x3 ru1x1b cc2 dse3x3 x2; do
  cap r02_conv_$c conv_tc_kernel 2 1 python tools/prof_conv.py $c planes
done
rm -f gpurun_out/r02_conv_*.ncu-rep
ls -la gpurun_out/r02_* | cut -c1-150
