#!/usr/bin/env bash
# GPU call r3a (2 GPUs): NVLink byte counters against a direct count
O=gpurun_out/r3a; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
CT_STATS=1 $TR tools/check_traffic.py > $O/traffic_stats.log 2>&1; echo rc=$?; tail -1 $O/traffic_stats.log
CT_STEPS=30 $TR tools/check_traffic.py > $O/traffic_30.log 2>&1; echo rc=$?; tail -1 $O/traffic_30.log
CT_P=3000000 $TR tools/check_traffic.py > $O/traffic_3m.log 2>&1; echo rc=$?; tail -1 $O/traffic_3m.log
