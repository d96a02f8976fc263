#!/usr/bin/env bash
# GPU call r2p: ncu of the two preprocess kernels (cfg2, cfg5)
O=gpurun_out/r2p; mkdir -p $O
for cfg in cfg2 cfg5; do
python tools/stage_times.py --cfg $cfg --iters 5 > $O/plain_$cfg.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:precull_kernel|preprocess_kernel" -s 10 -c 2 -o $O/pre_$cfg python tools/stage_times.py --cfg $cfg --iters 5 > $O/ncu_$cfg.log 2>&1
echo "$cfg rc=$?"
done
ls -la $O
