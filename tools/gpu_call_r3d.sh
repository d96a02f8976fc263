#!/usr/bin/env bash
O=gpurun_out/r3d; mkdir -p $O
python -m pytest tests/test_gpu_kernels.py -x -q -k "peer_exchange" > $O/pytest.log 2>&1; tail -30 $O/pytest.log | cut -c1-300
