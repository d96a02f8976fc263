#!/usr/bin/env bash
# GPU call (1 GPU): the round-end sequence without the ncu passes
O=gpurun_out/final_lite; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
python bench.py > $O/bench1.json 2> $O/bench1.err; echo "bench rc=$?"
for cfg in cfg1 cfg4; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg5 --no-shim --config $cfg > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "$cfg rc=$?"; done
