#!/usr/bin/env bash
# GPU call (2 GPUs): C mapper vs NCCL vs sequential, Python peer path, bench at N=2 (small, then full)
O=gpurun_out/checks2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > $O/topo.txt 2>&1
$TR tools/check_cmapper.py > $O/check_cmapper.json 2> $O/check_cmapper.err; echo "check_cmapper rc=$?"; cat $O/check_cmapper.json; tail -3 $O/check_cmapper.err
$TR tools/check_peer.py > $O/check_peer.log 2>&1; echo "check_peer rc=$?"; tail -8 $O/check_peer.log
$TR bench.py --gpus 2 --steps 10 --warmup 3 --config small > $O/bench2_small.json 2> $O/bench2_small.err; echo "bench small rc=$?"; tail -3 $O/bench2_small.err
python -c "
import json; d=json.load(open('$O/bench2_small.json')); print(json.dumps({k:d.get(k) for k in ('value','e2e','exchange_parity','nvlink')})[:1500]); print(json.dumps(d['mapping_cfg5'])[:1200])"
$TR bench.py --gpus 2 --steps 30 --warmup 5 > $O/bench2.json 2> $O/bench2.err; echo "bench rc=$?"; tail -3 $O/bench2.err
python -c "
import json; d=json.load(open('$O/bench2.json')); print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','e2e','exchange_parity','nvlink','rank_breakdown')})[:2500]); print(json.dumps(d['mapping_cfg5'])[:2500])"
python -m pytest tests/test_gpu_parity.py -x -q -k "grad_rows or accumulate" 2>&1 | tail -3
