#!/bin/bash
# round-2 evidence (run on the B200 box): launch list of the bench command, full ncu captures of the top kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --legs none --no-cpu-baseline --no-e2e --no-clocks"
$CMD > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err || { tail -5 gpurun_out/r02_bench_plain.err; exit 1; }
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_launches.csv $CMD > /dev/null 2>&1
HOT="python bench.py --steps 2 --warmup 3 --legs hotpath --no-cpu-baseline --no-e2e --no-clocks"
# the hot-path-only leg launches the same kernels on the same shapes back to back: capture the top kernels there
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"mwa_sp_kernel" -c 2 -o gpurun_out/r02_attn_sp -f python tools/prof_one.py attn8 auto blob > gpurun_out/r02_ncu_attn.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"mwa_sp_kernel" -c 1 -o gpurun_out/r02_attn_sp_dense -f python tools/prof_one.py attn8 auto ones > gpurun_out/r02_ncu_attn_dense.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"gdn_tc" -c 6 -o gpurun_out/r02_gdn_tc -f python tools/prof_gdn_sites.py > gpurun_out/r02_ncu_gdn.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"mwa_small_kernel" -c 1 -o gpurun_out/r02_attn_small -f python tools/prof_one.py attn4 auto ones > gpurun_out/r02_ncu_small.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"mwa_copy_dropped" -c 1 -o gpurun_out/r02_copy_dropped -f python tools/prof_one.py attn8 auto blob > gpurun_out/r02_ncu_copy.log 2>&1
for n in r02_attn_sp r02_attn_sp_dense r02_gdn_tc r02_attn_small r02_copy_dropped; do
  ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/${n}_ncu_raw.csv 2>/dev/null
done
ls -la gpurun_out/r02_* | cut -c1-150
