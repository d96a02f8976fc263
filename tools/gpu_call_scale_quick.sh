#!/usr/bin/env bash
# GPU call (N GPUs): headline loop only (no cfg5 block, no parity block) — e2e experiments
N=${1:-8}; O=gpurun_out/scale_quick; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
$TR bench.py --gpus $N --steps 40 --warmup 5 --no-cfg5 --no-parity > $O/bench$N.json 2> $O/bench$N.err; echo "bench rc=$?"; tail -3 $O/bench$N.err
python -c "
import json; d=json.load(open('$O/bench$N.json')); print(d['value'], d['ms_per_step'], json.dumps(d["e2e"]), d.get("rank_breakdown",[None])[0])"
nproc; cat /proc/cpuinfo | grep "model name" | sort | uniq -c | head -2
