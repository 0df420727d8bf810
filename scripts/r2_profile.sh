#!/bin/bash
# GPU evidence pass of round 2: in-graph timeline, ncu --set full of the decode kernels, format table.  Outputs -> gpurun_out/
# (the .ncu-rep files are exported to CSV on the box and removed: gpurun_out/ only travels back below 64 MiB)
mkdir -p gpurun_out
TIMELINE=2000 TL_ROWS=70 timeout 300 python scripts/probe_model.py "dict()" "dict(mega=1)" > gpurun_out/r2_timeline_q8_0.log 2>&1
WTYPE=q4_0 TIMELINE=2000 TL_ROWS=70 timeout 300 python scripts/probe_model.py > gpurun_out/r2_timeline_q4_0.log 2>&1
for spec in "q8_0 4096 14336 2 1 w13" "q8_0 14336 4096 1 0 w2" "q4_0 4096 14336 2 1 w13"; do
  set -- $spec
  timeout 200 python scripts/kernel_only.py $1 $2 $3 $4 $5 > gpurun_out/r2_kernel_$1_$6.txt 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:matvec_idp -s 12 -c 1 -f -o /tmp/r2_ncu_$1_$6 python scripts/kernel_only.py $1 $2 $3 $4 $5 > gpurun_out/r2_ncu_$1_$6.log 2>&1
  ncu -i /tmp/r2_ncu_$1_$6.ncu-rep --page raw --csv > gpurun_out/r2_ncu_$1_$6_raw.csv 2>/dev/null
  ncu -i /tmp/r2_ncu_$1_$6.ncu-rep --page source --csv > gpurun_out/r2_ncu_$1_$6_source.csv 2>/dev/null
  ncu -i /tmp/r2_ncu_$1_$6.ncu-rep --page details > gpurun_out/r2_ncu_$1_$6_details.txt 2>/dev/null
done
timeout 600 python scripts/format_table.py f16 bf16 f8_e4m3 q8_0 q5_1 q5_0 q4_1 q4_0 > gpurun_out/r2_format_table.md 2>&1
du -sh gpurun_out
tail -3 gpurun_out/r2_format_table.md
