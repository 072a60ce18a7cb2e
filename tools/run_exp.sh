#!/bin/bash
# one-off experiments
tag=${1:-exp}
out=gpurun_out
mkdir -p $out
set -x
python -m pytest tests/test_csr_builder.py -m gpu -x -q > $out/pytest_k4_$tag.log 2>&1; echo "pytest exit $?"; tail -2 $out/pytest_k4_$tag.log
python tools/bench_builder.py --config c3q > $out/builder_c3q_$tag.log 2>&1; tail -5 $out/builder_c3q_$tag.log
python tools/sweep.py --config c3q --mode exact --reps 3 --grid '[{},{"cta_capacity":64},{"cta_capacity":64,"cta2_capacity":8192}]' --forest-cache /tmp/c3q.npz > $out/sweep_c3q_exact_$tag.log 2>&1; tail -5 $out/sweep_c3q_exact_$tag.log | cut -c1-220
