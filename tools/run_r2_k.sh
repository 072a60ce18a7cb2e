#!/bin/bash
# Round 2, pass K: select4 v2 (de-unrolled, 256-bit loads, per-row window, reverse second pass, row-major K1 emit).
tag=${1:-r2_k}
out=gpurun_out
mkdir -p $out
set -x
timeout 900 python -m pytest tests/test_select3.py -x -q > $out/pytest_select34_$tag.log 2>&1; echo "select3/4 tests exit $?"; tail -15 $out/pytest_select34_$tag.log | cut -c1-300
timeout 900 python tools/fuzz_parity.py --cases 16 --seed 11 > $out/fuzz_$tag.log 2>&1; echo "fuzz exit $?"; tail -4 $out/fuzz_$tag.log | cut -c1-600
B="python bench.py --no-cpu-baseline --single-mode --no-estimator-e2e --steps 10"
timeout 1500 $B --option select_impl=4 > $out/bench_c3_sel4_c4_$tag.json 2> $out/bench_c3_sel4_c4_$tag.err; echo "sel4 exit $?"; tail -3 $out/bench_c3_sel4_c4_$tag.err
timeout 900 $B --option select_impl=4 --option select4_ctas=3 > $out/bench_c3_sel4_c3_$tag.json 2> $out/bench_c3_sel4_c3_$tag.err; echo "sel4 3 ctas exit $?"
timeout 900 $B --option select_impl=4 --option rows_per_chunk=1048576 > $out/bench_c3_sel4_c4_chunk1m_$tag.json 2> $out/bench_c3_sel4_c4_chunk1m_$tag.err; echo "sel4 1M exit $?"
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
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"quantile_select4|apply_kernel" -s 8 -c 2 -f -o $out/prof_c3_sel4_$tag $B2 > $out/ncu_$tag.log 2>&1; echo "ncu exit $?"
ls -la $out/*$tag*
du -sh $out
