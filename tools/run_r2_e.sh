#!/bin/bash
# Round 2, pass E: warp-specialised select3: tests, C3 bench (fast only), ncu of one step.
tag=${1:-r2_e}
out=gpurun_out
mkdir -p $out
set -x
timeout 900 python -m pytest tests/test_select3.py -x -q > $out/pytest_select3_$tag.log 2>&1; echo "select3 tests exit $?"; tail -15 $out/pytest_select3_$tag.log
timeout 1500 python bench.py --no-cpu-baseline --single-mode > $out/bench_c3_$tag.json 2> $out/bench_c3_$tag.err; echo "bench exit $?"; tail -5 $out/bench_c3_$tag.err; python - <<PY
import json
try:
    d=json.loads(open('$out/bench_c3_$tag.json').read().strip().splitlines()[-1])
    print(d['ms_per_step'], d['kernel_ms'], d['e2e']['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['parity_sample'], d.get('other_mode'))
except Exception as e: print('no bench line', e)
PY
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-estimator-e2e --single-mode"
timeout 900 $B --mode fast > $out/plain_fast_$tag.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"quantile_select3" -s 4 -c 1 -f -o $out/prof_c3_fast_$tag \
    $B --mode fast > $out/ncu_fast_$tag.log 2>&1; echo "ncu fast exit $?"
ls -la $out/*$tag*
