#!/usr/bin/env bash
# GPU call r2e: r1 blend + r1 preprocess with hit list + distribute; full gpu suite; stage times; bench
O=gpurun_out/r2e; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
for cfg in cfg2 cfg1 cfg4 cfg5; do
  python tools/stage_times.py --cfg $cfg --tag new >> $O/stage_times.jsonl 2>>$O/stage_err_new.log
  GSR_LIB_PATH=$PWD/variants/libgsr_r1.so python tools/stage_times.py --cfg $cfg --tag r1 >> $O/stage_times.jsonl 2>>$O/stage_err_r1.log
done
cat $O/stage_times.jsonl
python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print(d['value'], d['e2e']['value'], d.get('rasterizer_only',{}).get('value'), d['mapping_cfg5']['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()})"
