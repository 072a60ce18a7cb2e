#!/bin/bash
# select2 iteration at full C3 scale (DRAM-resident members): parity tests, C3-quarter timing, full C3 fast bench.
tag=${1:-c3f}
out=gpurun_out
mkdir -p $out
set -x
python -m pytest tests -m gpu -x -q -k "fast or golden" > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -2 $out/pytest_gpu_$tag.log
python tools/sweep.py --config c3q --mode fast --reps 3 --grid '[{}]' --forest-cache /tmp/c3q.npz > $out/sweep_c3q_$tag.log 2>&1; tail -2 $out/sweep_c3q_$tag.log | cut -c1-200
timeout 1500 python bench.py --config c3 --steps 5 --warmup 3 --mode fast --no-cpu-baseline --forest-cache /tmp/c3.npz > $out/bench_c3_fast_$tag.json 2> $out/bench_c3_fast_$tag.err; echo "fast exit $?"; python - <<PY
import json
d=json.loads(open('$out/bench_c3_fast_$tag.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['kernel_ms'], d['e2e']['ms_per_step'])
PY
