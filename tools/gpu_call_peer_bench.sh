#!/usr/bin/env bash
# GPU call: local efficiency of the peer kernel (n_peers = 1) + ncu
O=gpurun_out/peer_bench; mkdir -p $O
python tools/peer_adam_bench.py 3000000 0.15 > $O/peer_bench.log 2>&1; cat $O/peer_bench.log
ncu --set full --clock-control none --import-source on -k regex:peer_adam_kernel -s 14 -c 1 -o $O/peer python tools/peer_adam_bench.py 3000000 0.15 > $O/ncu.log 2>&1; echo "ncu rc=$?"
