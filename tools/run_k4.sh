#!/bin/bash
# K4 pass: parity tests (all), builder timings on C2 and C3-quarter.
tag=${1:-k4}
out=gpurun_out
mkdir -p $out
set -x
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -5 $out/pytest_gpu_$tag.log
python tools/bench_builder.py --config c2 > $out/builder_c2_$tag.log 2>&1; tail -5 $out/builder_c2_$tag.log
python tools/bench_builder.py --config c3q > $out/builder_c3q_$tag.log 2>&1; tail -5 $out/builder_c3q_$tag.log
python tools/sweep.py --config c2 --reps 5 --grid '[{},{"warp_unit":0}]' --forest-cache /tmp/c2.npz > $out/sweep_c2_$tag.log 2>&1; tail -3 $out/sweep_c2_$tag.log | cut -c1-200
