#!/usr/bin/env bash
# GPU call r2c: persistent blend (in place) vs the round-1 library, full gpu suite, bench (N=1)
O=gpurun_out/r2c; mkdir -p $O
python tools/stage_times.py --cfg cfg2 --tag new >> $O/stage_times.jsonl 2>$O/stage_err_new.log
GSR_LIB_PATH=$PWD/variants/libgsr_r1.so python tools/stage_times.py --cfg cfg2 --tag r1 >> $O/stage_times.jsonl 2>$O/stage_err_r1.log
GSR_LIB_PATH=$PWD/variants/libgsr_b.so python tools/stage_times.py --cfg cfg2 --tag b >> $O/stage_times.jsonl 2>$O/stage_err_b.log
cat $O/stage_times.jsonl
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
for cfg in cfg1 cfg4; do python tools/stage_times.py --cfg $cfg --tag new >> $O/stage_times.jsonl 2>>$O/stage_err_new.log; GSR_LIB_PATH=$PWD/variants/libgsr_r1.so python tools/stage_times.py --cfg $cfg --tag r1 >> $O/stage_times.jsonl 2>>$O/stage_err_r1.log; done
python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
python bench.py --impl reference --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_ref.json 2> $O/bench_ref.err; tail -3 $O/bench_ref.err
tail -4 $O/stage_times.jsonl
