#!/usr/bin/env bash
# GPU call r2b: persistent blend variants (timing + parity), drop-in consumer test, loss oracle test
O=gpurun_out/r2b; mkdir -p $O
for v in "" a b; do
  lib=""; [ -n "$v" ] && lib="$PWD/variants/libgsr_$v.so"
  GSR_LIB_PATH=${lib:-$PWD/gsdf_slam_b200/libgsr.so} python tools/stage_times.py --cfg cfg2 --tag "v$v" >> $O/stage_times.jsonl 2>$O/stage_err_$v.log
done
for v in a b; do
  GSR_LIB_PATH=$PWD/variants/libgsr_$v.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_parity_full.py -x -q > $O/parity_$v.log 2>&1
  tail -3 $O/parity_$v.log
done
for v in a; do
  for cfg in cfg1 cfg4; do GSR_LIB_PATH=$PWD/variants/libgsr_$v.so python tools/stage_times.py --cfg $cfg --tag "v$v" >> $O/stage_times.jsonl 2>>$O/stage_err_$v.log; done
done
for cfg in cfg1 cfg4; do python tools/stage_times.py --cfg $cfg --tag "v" >> $O/stage_times.jsonl 2>>$O/stage_err_.log; done
python -m pytest tests/test_dropin_consumer.py tests/test_gpu_loss.py -x -q > $O/dropin_loss.log 2>&1; tail -3 $O/dropin_loss.log
cat $O/stage_times.jsonl
