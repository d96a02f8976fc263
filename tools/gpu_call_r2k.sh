#!/usr/bin/env bash
# GPU call r2k: ncu evidence of round 2 — launch list, full-set capture of one step's kernels (cfg2), DRAM bytes per config
O=gpurun_out/r2k; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cfg5 --no-shim"
KRE='regex:preprocess_kernel|tile_meta_kernel|emit_kernel|emit_large_kernel|tile_sort_pack|blend_fwd_kernel|blend_bwd_kernel|preprocess_bwd_kernel|adam_rows_all_kernel'
$CMD > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
echo "launches rc=$?"
$CMD > $O/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "$KRE" -s 45 -c 10 -o $O/top $CMD > $O/ncu_full.log 2>&1
echo "full rc=$?"; tail -2 $O/ncu_full.log
for cfg in cfg1 cfg4; do
  $CMD --config $cfg > $O/plain_$cfg.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KRE" -s 45 -c 12 --csv --log-file $O/dram_$cfg.csv $CMD --config $cfg > $O/ncu_$cfg.log 2>&1
  echo "$cfg rc=$?"
done
ls -la $O
