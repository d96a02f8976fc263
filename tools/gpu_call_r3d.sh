#!/usr/bin/env bash
O=gpurun_out/r3d; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "grad_rows" > $O/pytest.log 2>&1; tail -30 $O/pytest.log | cut -c1-400
