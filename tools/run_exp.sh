#!/bin/bash
# one-off experiments
tag=${1:-exp}
out=gpurun_out
mkdir -p $out
set -x
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest_gpu_$tag.log
python bench.py > $out/bench_c2_$tag.json 2> $out/bench_c2_$tag.err; echo "bench exit $?"; cut -c1-600 $out/bench_c2_$tag.json
python tools/sweep.py --config c3q --mode fast --reps 5 --grid "[{}]" --forest-cache /tmp/c3q.npz > $out/sweep_c3q_$tag.log 2>&1; tail -2 $out/sweep_c3q_$tag.log | cut -c1-200
