#!/usr/bin/env bash
# GPU call r2j: packed (FFMA2) blend kernels vs scalar: timing and parity
O=gpurun_out/r2j; mkdir -p $O
for cfg in cfg2 cfg1 cfg5; do for pk in 0 1; do
  GSR_BLEND_PACKED=$pk python tools/stage_times.py --cfg $cfg --tag packed$pk >> $O/stage_times.jsonl 2>>$O/stage_err.log
done; done
cat $O/stage_times.jsonl
GSR_BLEND_PACKED=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_parity_full.py -x -q > $O/parity_packed.log 2>&1; tail -6 $O/parity_packed.log
