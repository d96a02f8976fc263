#!/usr/bin/env bash
# GPU call r2m: two-stage preprocess (conservative cull + candidate kernel): parity, stage times vs the one-kernel build
O=gpurun_out/r2m; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
for cfg in cfg2 cfg1 cfg4 cfg5; do
  python tools/stage_times.py --cfg $cfg --tag split >> $O/stage_times.jsonl 2>>$O/stage_err_new.log
  GSR_LIB_PATH=$PWD/variants/libgsr_r2base.so python tools/stage_times.py --cfg $cfg --tag base >> $O/stage_times.jsonl 2>>$O/stage_err_base.log
done
cut -c1-400 $O/stage_times.jsonl
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print(d['value'], d['e2e']['value'], d.get('rasterizer_only',{}).get('value'), d['mapping_cfg5']['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()})"
