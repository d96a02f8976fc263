#!/usr/bin/env bash
# GPU call r3e (8 GPUs): bench at N=8 (headline weak scaling + mapping_cfg5 strong scaling + exchange parity)
O=gpurun_out/r3e; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
$TR bench.py --gpus 8 --steps 30 --warmup 5 > $O/bench8.json 2> $O/bench8.err; echo "bench rc=$?"; tail -3 $O/bench8.err
python -c "
import json; d=json.load(open('$O/bench8.json')); print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','e2e','exchange_parity','nvlink')})[:3000]); print(json.dumps(d.get('rank_breakdown'))[:1500]); print(json.dumps(d['mapping_cfg5'])[:3000])"
