#!/bin/bash
# Round 2, pass N: select4 v4 (predicated octet loads, one-pass list build, rank-interval hit test) at C3.
tag=${1:-r2_n}
out=gpurun_out
mkdir -p $out
set -x
for c in c1_etqr_boston rf_pure_leaves et_depth2_large_sets; do timeout 120 python tools/dbg_select4.py $c 2>&1 | tail -2 | cut -c1-160; done
timeout 900 python -m pytest tests/test_select3.py -x -q > $out/pytest_select34_$tag.log 2>&1; echo "select3/4 tests exit $?"; tail -5 $out/pytest_select34_$tag.log | cut -c1-300
timeout 600 python tools/fuzz_parity.py --cases 12 --seed 11 > $out/fuzz_$tag.log 2>&1; echo "fuzz exit $?"; tail -2 $out/fuzz_$tag.log | cut -c1-300
B="python bench.py --no-cpu-baseline --single-mode --no-estimator-e2e --steps 10"
for u in 2 4; do for c in 4 2; do
timeout 1500 $B --option select_impl=4 --option select4_unroll=$u --option select4_ctas=$c > $out/bench_c3_sel4_u${u}_c${c}_$tag.json 2> $out/bench_c3_sel4_u${u}_c${c}_$tag.err; echo "sel4 unroll $u ctas $c exit $?"
done; done
python - <<PY
import json, glob
for f in sorted(glob.glob('$out/bench_c3_*_$tag.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['ms_per_step'], d['kernel_ms'], d['e2e']['ms_per_step'], d['roofline']['kernel'], d['parity_sample']['ok'], d['parity_sample']['max_abs_err'])
    except Exception as e: print('no bench line', f, e)
PY
B2="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-estimator-e2e --single-mode --option select_impl=4"
timeout 900 $B2 > $out/plain_$tag.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"quantile_select4" -s 4 -c 1 -f -o $out/prof_c3_sel4_v4_$tag $B2 > $out/ncu_$tag.log 2>&1; echo "ncu exit $?"
ls -la $out/*$tag*
du -sh $out
