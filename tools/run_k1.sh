#!/bin/bash
# K1 (leaf lookup) iteration pass: parity tests, narrow vs two-level records at several ILP on C2 and C3-quarter, e2e pipeline, ncu.
tag=${1:-k1}
out=gpurun_out
mkdir -p $out
set -x
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -5 $out/pytest_gpu_$tag.log
G='[{"apply_wide":0},{"apply_wide":1,"apply_ilp":1},{"apply_wide":1,"apply_ilp":2},{"apply_wide":1,"apply_ilp":8},{"apply_wide":1,"apply_ilp":4},{"apply_wide":1,"apply_ilp":2,"apply_row_warps":2}]'
python tools/sweep.py --config c2 --reps 5 --grid "$G" --forest-cache /tmp/c2.npz > $out/sweep_c2_$tag.log 2>&1; tail -8 $out/sweep_c2_$tag.log | cut -c1-220
python tools/e2e_sweep.py --config c2 --forest-cache /tmp/c2.npz > $out/e2e_c2_$tag.log 2>&1; tail -7 $out/e2e_c2_$tag.log
python tools/sweep.py --config c3q --mode fast --reps 3 --grid "$G" --forest-cache /tmp/c3q.npz > $out/sweep_c3q_$tag.log 2>&1; tail -8 $out/sweep_c3q_$tag.log | cut -c1-220
ncu --set full --clock-control none --import-source on -k regex:"apply_wide_kernel|quantile_select2" -s 2 -c 2 -f -o $out/prof_c3q_$tag \
    python tools/sweep.py --config c3q --mode fast --reps 1 --rows 65536 --grid '[{}]' --forest-cache /tmp/c3q.npz > $out/ncu_c3q_$tag.log 2>&1; echo "ncu exit $?"
