#!/usr/bin/env bash
# GPU call r2y: ncu of preprocess_bwd (cfg5, cfg2)
O=gpurun_out/r2y; mkdir -p $O
for cfg in cfg5 cfg2; do
python tools/stage_times.py --cfg $cfg --iters 5 > $O/plain_$cfg.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:preprocess_bwd_kernel" -s 6 -c 1 -o $O/pb_$cfg python tools/stage_times.py --cfg $cfg --iters 5 > $O/ncu_$cfg.log 2>&1
echo "$cfg rc=$?"
done
