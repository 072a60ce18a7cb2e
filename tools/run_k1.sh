#!/bin/bash
# K1 (leaf lookup) iteration pass: parity tests, ILP / row_warps sweep on C2 and C3-quarter, ncu of apply on C3-quarter.
tag=${1:-k1}
out=gpurun_out
mkdir -p $out
set -x
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest_gpu_$tag.log
G='[{"apply_ilp":2},{"apply_ilp":4},{"apply_ilp":8},{"apply_ilp":4,"apply_row_warps":1},{"apply_ilp":4,"apply_row_warps":2},{"apply_ilp":8,"apply_row_warps":1},{"apply_ilp":2,"apply_row_warps":1}]'
python tools/sweep.py --config c2 --reps 5 --grid "$G" --forest-cache /tmp/c2.npz > $out/sweep_c2_$tag.log 2>&1; tail -9 $out/sweep_c2_$tag.log
python tools/sweep.py --config c3q --mode fast --reps 3 --grid "$G" --forest-cache /tmp/c3q.npz > $out/sweep_c3q_$tag.log 2>&1; tail -9 $out/sweep_c3q_$tag.log
ncu --set full --clock-control none --import-source on -k regex:"apply_kernel" -s 1 -c 1 -f -o $out/prof_apply_c3q_$tag \
    python tools/sweep.py --config c3q --mode fast --reps 1 --rows 65536 --grid '[{}]' --forest-cache /tmp/c3q.npz > $out/ncu_apply_c3q_$tag.log 2>&1; echo "ncu exit $?"
