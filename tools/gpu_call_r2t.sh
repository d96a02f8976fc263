#!/usr/bin/env bash
# GPU call r2t: two-stage preprocess, second version: parity + stage times (32 / 16 / 8 candidates per warp, one-kernel build)
O=gpurun_out/r2t; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_parity_full.py tests/test_gpu_raw_params.py -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
for cfg in cfg2 cfg5 cfg1 cfg4; do
  python tools/stage_times.py --cfg $cfg --tag cq512 >> $O/stage_times.jsonl 2>>$O/stage_err.log
  for v in r2base; do
    GSR_LIB_PATH=$PWD/variants/libgsr_$v.so python tools/stage_times.py --cfg $cfg --tag $v >> $O/stage_times.jsonl 2>>$O/stage_err.log
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2t/stage_times.jsonl'):
    d=json.loads(l); s=d['stages_us']; print(d['cfg'], d['tag'], d['fwd_bwd_ms'], 'precull', s.get('precull'), 'pre', s['preprocess'], 'emit', s['emit'])
PY
