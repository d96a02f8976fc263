#!/usr/bin/env bash
# GPU call r2n: cull / candidate kernel times separately
O=gpurun_out/r2n; mkdir -p $O
for cfg in cfg2 cfg5 cfg4; do
  python tools/stage_times.py --cfg $cfg --tag split >> $O/stage_times.jsonl 2>>$O/stage_err_new.log
done
cut -c1-500 $O/stage_times.jsonl
