#!/usr/bin/env bash
# GPU call (1 GPU): the round-end sequence + ncu evidence of the final kernels
O=gpurun_out/final1; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
python bench.py > $O/bench1.json 2> $O/bench1.err; echo "bench rc=$?"
for cfg in cfg1 cfg4; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg5 --no-shim --config $cfg > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "$cfg rc=$?"; done
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cfg5 --no-shim"
KRE='regex:precull_kernel|preprocess_kernel|count_large_kernel|tile_meta_kernel|emit_kernel|emit_large_kernel|tile_sort_pack|tile_select_kernel|blend_fwd_kernel|blend_bwd_kernel|select_grad_rows_kernel|preprocess_bwd_kernel|adam_rows_all_kernel'
$CMD > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
echo "launches rc=$?"
$CMD > $O/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "$KRE" -s 56 -c 14 -o $O/top $CMD > $O/ncu_full.log 2>&1
echo "full rc=$?"; tail -2 $O/ncu_full.log
for cfg in cfg1 cfg4; do
  $CMD --config $cfg > $O/plain_$cfg.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KRE" -s 56 -c 16 --csv --log-file $O/dram_$cfg.csv $CMD --config $cfg > $O/ncu_$cfg.log 2>&1
  echo "$cfg rc=$?"
done
ls -la $O | head -30
