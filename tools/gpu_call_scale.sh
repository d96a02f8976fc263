#!/usr/bin/env bash
# GPU call (N GPUs): bench.py at N ranks (headline weak scaling + mapping_cfg5 strong scaling + exchange parity)
N=${1:-8}; O=gpurun_out/scale_r2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
$TR bench.py --gpus $N --steps 30 --warmup 5 > $O/bench$N.json 2> $O/bench$N.err; echo "bench rc=$?"; tail -3 $O/bench$N.err
python -c "
import json; d=json.load(open('$O/bench$N.json')); print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','e2e','nvlink')})[:1500]); print(json.dumps(d['exchange_parity'])[:1200]); print(json.dumps(d.get('rank_breakdown'))[:700]); c=d['mapping_cfg5']; print(c['ms_per_step'], json.dumps(c.get('rank_breakdown'))[:700], json.dumps(c.get('nvlink'))[:600])"
