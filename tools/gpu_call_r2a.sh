#!/usr/bin/env bash
# GPU call r2a: micro-benchmark, occupancy knob, full-size parity, baseline bench
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
./variants/issue_bench > $O/issue_bench.txt 2>&1
for pad in 0 9000 14000; do
  GSR_BLEND_SMEM_PAD=$pad python tools/stage_times.py --cfg cfg2 --tag pad$pad >> $O/stage_times.jsonl 2>$O/stage_err_$pad.log
done
python -m pytest tests/test_gpu_parity_full.py -x -q -s > $O/full_parity.log 2>&1
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
tail -5 $O/full_parity.log; cat $O/stage_times.jsonl; head -40 $O/issue_bench.txt
