#!/usr/bin/env bash
# GPU call r2v: the round-end sequence on one GPU — full -m gpu suite, smoke, both bench arms
O=gpurun_out/r2v; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
python bench.py > $O/bench1.json 2> $O/bench1.err; echo "bench rc=$?"
cut -c1-1500 $O/bench1.json
