#!/bin/bash
# one-off experiments
tag=${1:-exp}
out=gpurun_out
mkdir -p $out
set -x
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest_gpu_$tag.log
G='[{"apply_stage":0},{"apply_stage":1},{"apply_stage":1,"apply_ilp":2},{"apply_stage":1,"apply_ilp":8}]'
python tools/sweep.py --config c2 --reps 7 --grid "$G" --forest-cache /tmp/c2.npz > $out/sweep_c2_$tag.log 2>&1; tail -5 $out/sweep_c2_$tag.log | cut -c1-200
python tools/sweep.py --config c3q --mode fast --reps 3 --grid "$G" --forest-cache /tmp/c3q.npz > $out/sweep_c3q_$tag.log 2>&1; tail -5 $out/sweep_c3q_$tag.log | cut -c1-200
