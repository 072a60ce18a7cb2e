#!/bin/bash
# one-off experiment: leaf lookup compiled for 8 CTAs per SM (32 registers)
tag=${1:-exp}
out=gpurun_out
mkdir -p $out
python tools/sweep.py --config c2 --reps 9 --grid '[{},{"apply_levels":4}]' > $out/sweep_c2_$tag.log 2>&1; grep "apply=" $out/sweep_c2_$tag.log | cut -c1-150
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -2 $out/pytest_gpu_$tag.log
