#!/bin/bash
# one-off experiments
tag=${1:-exp}
out=gpurun_out
mkdir -p $out
set -x
G='[{},{"apply_ilp":8},{"apply_ilp":2},{"apply_row_warps":2},{"apply_row_warps":2,"apply_ilp":8}]'
python tools/sweep.py --config c2 --reps 7 --grid "$G" --forest-cache /tmp/c2.npz > $out/sweep_c2_$tag.log 2>&1; tail -7 $out/sweep_c2_$tag.log | cut -c1-200
python tools/sweep.py --config c3q --mode fast --reps 3 --grid "$G" --forest-cache /tmp/c3q.npz > $out/sweep_c3q_$tag.log 2>&1; tail -7 $out/sweep_c3q_$tag.log | cut -c1-200
