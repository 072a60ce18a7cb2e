#!/bin/bash
# Round 2, pass O: select4 v4 with / without L2 eviction hints, 3 / 2 / 1 CTAs per SM, at C3.
tag=${1:-r2_o}
out=gpurun_out
mkdir -p $out
set -x
timeout 900 python -m pytest tests/test_select3.py -x -q -k select4 > $out/pytest_select4_$tag.log 2>&1; echo "select4 tests exit $?"; tail -3 $out/pytest_select4_$tag.log | cut -c1-300
B="python bench.py --no-cpu-baseline --single-mode --no-estimator-e2e --steps 10"
for h in 1 0; do for c in 4 2 1; do
timeout 1500 $B --option select_impl=4 --option select4_hints=$h --option select4_ctas=$c > $out/bench_c3_sel4_h${h}_c${c}_$tag.json 2> $out/bench_c3_sel4_h${h}_c${c}_$tag.err; echo "sel4 hints $h ctas $c exit $?"
done; done
timeout 900 $B --option select_impl=4 --option select4_hints=1 --option rows_per_chunk=1048576 > $out/bench_c3_sel4_h1_chunk1m_$tag.json 2> $out/bench_c3_sel4_h1_chunk1m_$tag.err; echo "sel4 1M exit $?"
timeout 900 $B > $out/bench_c3_sel2_$tag.json 2> $out/bench_c3_sel2_$tag.err; echo "sel2 exit $?"
python - <<PY
import json, glob
for f in sorted(glob.glob('$out/bench_c3_*_$tag.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['ms_per_step'], d['kernel_ms'], d['e2e']['ms_per_step'], d['roofline']['kernel'], d['parity_sample']['ok'], d['parity_sample']['max_abs_err'])
    except Exception as e: print('no bench line', f, e)
PY
B2="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-estimator-e2e --single-mode --option select_impl=4 --option select4_hints=1"
timeout 900 $B2 > $out/plain_$tag.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"quantile_select4" -s 4 -c 1 -f -o $out/prof_c3_sel4_v4h_$tag $B2 > $out/ncu_$tag.log 2>&1; echo "ncu exit $?"
ls -la $out/*$tag*
du -sh $out
