#!/usr/bin/env bash
# GPU call (1 GPU, development): parity suites + stage times of the four configs
O=gpurun_out/dev; mkdir -p $O; rm -f $O/stage_times.jsonl
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_parity_full.py -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log | cut -c1-300
for cfg in cfg2 cfg5 cfg1 cfg4; do python tools/stage_times.py --cfg $cfg --iters 20 --tag new >> $O/stage_times.jsonl 2>>$O/stage_err.log; done
python - <<'PY'
import json
for l in open('gpurun_out/dev/stage_times.jsonl'):
    d=json.loads(l); s=d['stages_us']; print(d['cfg'], d['tag'], d['fwd_bwd_ms'], {k:round(v,1) for k,v in s.items()})
PY
