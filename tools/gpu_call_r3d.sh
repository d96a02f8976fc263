#!/usr/bin/env bash
O=gpurun_out/r3d; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_parity_full.py -x -q > $O/pytest.log 2>&1; tail -8 $O/pytest.log | cut -c1-300
for cfg in cfg1 cfg4 cfg2; do python tools/stage_times.py --cfg $cfg --tag skipcull >> $O/stage_times.jsonl 2>>$O/stage_err.log; done
python - <<'PY'
import json
for l in open('gpurun_out/r3d/stage_times.jsonl'):
    d=json.loads(l); s=d['stages_us']; print(d['cfg'], d['tag'], d['fwd_bwd_ms'], 'precull', s.get('precull'), 'pre', s['preprocess'], 'emit', s['emit'], 'pre_bwd', s['preprocess_bwd'])
PY
