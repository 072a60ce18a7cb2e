#!/bin/bash
# one-off experiments: visits between two warp votes in the leaf lookup (apply_levels)
tag=${1:-exp}
out=gpurun_out
mkdir -p $out
G='[{},{"apply_levels":1},{"apply_levels":3},{"apply_levels":4}]'
python tools/sweep.py --config c2 --reps 9 --grid "$G" > $out/sweep_c2_$tag.log 2>&1; grep "apply=" $out/sweep_c2_$tag.log | cut -c1-150
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -2 $out/pytest_gpu_$tag.log
