#!/usr/bin/env bash
# GPU call r2i (2 GPUs): vectorised stats exchange + parity metrics
O=gpurun_out/r2i; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
$TR tools/check_cmapper.py > $O/check_cmapper.json 2> $O/check_cmapper.err; echo "check_cmapper rc=$?"; cat $O/check_cmapper.json
$TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench2.json 2> $O/bench2.err; echo "bench rc=$?"; tail -3 $O/bench2.err
python -c "
import json; d=json.load(open('$O/bench2.json')); print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','exchange_parity')})[:3000]); print(json.dumps(d['mapping_cfg5'].get('rank_breakdown'))[:1500], d['mapping_cfg5']['ms_per_step'])"
