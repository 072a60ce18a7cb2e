#!/bin/bash
# K2 iteration pass: parity tests, C2 timings (warp kernel), C3-quarter timings old vs new selection kernel, ncu of the new one.
tag=${1:-k2}
out=gpurun_out
mkdir -p $out
set -x
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -5 $out/pytest_gpu_$tag.log
python tools/sweep.py --config c2 --reps 5 --grid '[{}]' --forest-cache /tmp/c2.npz > $out/sweep_c2_$tag.log 2>&1; tail -2 $out/sweep_c2_$tag.log | cut -c1-200
python tools/e2e_sweep.py --config c2 --forest-cache /tmp/c2.npz > $out/e2e_c2_$tag.log 2>&1; tail -7 $out/e2e_c2_$tag.log
python tools/sweep.py --config c3q --mode fast --reps 3 --grid '[{"select_impl":1},{"select_impl":0}]' --forest-cache /tmp/c3q.npz > $out/sweep_c3q_$tag.log 2>&1; tail -4 $out/sweep_c3q_$tag.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:"quantile_select2" -s 1 -c 1 -f -o $out/prof_select2_c3q_$tag \
    python tools/sweep.py --config c3q --mode fast --reps 1 --rows 65536 --grid '[{}]' --forest-cache /tmp/c3q.npz > $out/ncu_select2_c3q_$tag.log 2>&1; echo "ncu exit $?"
