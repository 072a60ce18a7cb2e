#!/bin/bash
# Round 2, pass F: whole GPU suite, C3 bench in both modes (finer exact buckets), select2 vs select3 A/B.
tag=${1:-r2_f}
out=gpurun_out
mkdir -p $out
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -15 $out/pytest_gpu_$tag.log
timeout 1500 python bench.py --no-cpu-baseline > $out/bench_c3_$tag.json 2> $out/bench_c3_$tag.err; echo "bench exit $?"; tail -3 $out/bench_c3_$tag.err
timeout 900 python bench.py --no-cpu-baseline --single-mode --no-estimator-e2e --option select_impl=2 > $out/bench_c3_sel2_$tag.json 2> $out/bench_c3_sel2_$tag.err; echo "sel2 exit $?"
python - <<PY
import json
for f in ('bench_c3_$tag.json','bench_c3_sel2_$tag.json'):
    try:
        d=json.loads(open('$out/'+f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['kernel_ms'], d['e2e']['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['parity_sample']['ok'], d.get('other_mode'))
    except Exception as e: print('no bench line', f, e)
PY
ls -la $out/*$tag*
