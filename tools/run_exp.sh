#!/bin/bash
# one-off experiments: node block layouts (QF_NODE_BLOCK_LEVELS, csrc/pack_forest.cpp)
tag=${1:-exp}
out=gpurun_out
mkdir -p $out
for h in 0 3 2; do
  export QF_NODE_BLOCK_LEVELS=$h
  python tools/sweep.py --config c2 --reps 7 --grid "[{}]" > $out/sweep_c2_${tag}_h$h.log 2>&1; echo "c2 h=$h: $(grep -m1 apply= $out/sweep_c2_${tag}_h$h.log | cut -c60-200)"
done
for h in 0 3; do
  export QF_NODE_BLOCK_LEVELS=$h
  python tools/sweep.py --config c3q --mode fast --reps 3 --grid "[{}]" > $out/sweep_c3q_${tag}_h$h.log 2>&1; echo "c3q h=$h: $(grep -m1 apply= $out/sweep_c3q_${tag}_h$h.log | cut -c60-200)"
done
