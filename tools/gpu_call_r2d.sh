#!/usr/bin/env bash
# GPU call r2d: front-end v2 (funnel preprocess + distribute) parity and timings; blend variants
O=gpurun_out/r2d; mkdir -p $O
for v in fe h28 h p1b r1; do
  GSR_LIB_PATH=$PWD/variants/libgsr_$v.so python tools/stage_times.py --cfg cfg2 --tag $v >> $O/stage_times.jsonl 2>$O/stage_err_$v.log
done
cat $O/stage_times.jsonl
GSR_LIB_PATH=$PWD/variants/libgsr_fe.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_parity_full.py tests/test_gpu_raw_params.py -x -q > $O/parity_fe.log 2>&1; tail -15 $O/parity_fe.log
for cfg in cfg1 cfg4 cfg5; do for v in fe r1; do GSR_LIB_PATH=$PWD/variants/libgsr_$v.so python tools/stage_times.py --cfg $cfg --tag $v >> $O/stage_times2.jsonl 2>>$O/stage_err_$v.log; done; done
cat $O/stage_times2.jsonl
