#!/usr/bin/env bash
# GPU call r3b: grad_rows protocol (rows without gradient skipped);: preprocess_bwd with hoisted loads / prefetched accumulate rows, persistent emit: full suite, stage times, bench
O=gpurun_out/r3b; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
for cfg in cfg2 cfg5 cfg1 cfg4; do
  python tools/stage_times.py --cfg $cfg --tag new >> $O/stage_times.jsonl 2>>$O/stage_err.log
  for v in r2base; do
    GSR_LIB_PATH=$PWD/variants/libgsr_$v.so python tools/stage_times.py --cfg $cfg --tag $v >> $O/stage_times.jsonl 2>>$O/stage_err.log
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r3b/stage_times.jsonl'):
    d=json.loads(l); s=d['stages_us']; print(d['cfg'], d['tag'], d['fwd_bwd_ms'], 'precull', s.get('precull'), 'pre', s['preprocess'], 'emit', s['emit'], 'pre_bwd', s['preprocess_bwd'])
PY
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print(d['value'], d['e2e']['value'], d.get('rasterizer_only',{}).get('value'), d['mapping_cfg5']['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()}); print(d['mapping_cfg5'].get('stages_ms_per_step_rank0'))"
