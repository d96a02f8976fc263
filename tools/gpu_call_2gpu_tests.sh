#!/usr/bin/env bash
# GPU call (2 GPUs): the two-devices-in-one-process test + the 2-GPU subprocess test of the suite
O=gpurun_out/tests2; mkdir -p $O
python -m pytest tests/test_gpu_kernels.py -x -q -k "two_devices or two_gpus" > $O/pytest.log 2>&1; tail -15 $O/pytest.log | cut -c1-300
