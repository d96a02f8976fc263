#!/usr/bin/env bash
# GPU call r2u: ncu of the cull kernel (cfg2)
O=gpurun_out/r2u; mkdir -p $O
for cfg in cfg2; do
python tools/stage_times.py --cfg $cfg --iters 5 > $O/plain_$cfg.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:precull_kernel|preprocess_kernel|count_large" -s 9 -c 3 -o $O/pre_$cfg python tools/stage_times.py --cfg $cfg --iters 5 > $O/ncu_$cfg.log 2>&1
echo "$cfg rc=$?"
done
