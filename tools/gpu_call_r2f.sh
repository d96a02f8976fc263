#!/usr/bin/env bash
O=gpurun_out/r2f; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print(d['value'], d['e2e']['value'], d.get('rasterizer_only',{}).get('value'), d['mapping_cfg5']['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()})"
