#!/bin/bash
# Round 2, pass P (last GPU minutes): compact lists with multi-octet leaves on even octets (64-byte bursts),
# select4 + hints at C3 with and without, after the bit-equality tests.
tag=${1:-r2_p}
out=gpurun_out
mkdir -p $out
set -x
timeout 120 python -m pytest tests/test_select3.py -x -q -k "aligned" > $out/pytest_align_$tag.log 2>&1; echo "align tests exit $?"; tail -2 $out/pytest_align_$tag.log | cut -c1-200
B="python bench.py --no-cpu-baseline --single-mode --no-estimator-e2e --steps 10 --option select_impl=4 --option select4_hints=1"
timeout 400 $B --option compact_align=1 > $out/bench_c3_sel4_align1_$tag.json 2> $out/bench_c3_sel4_align1_$tag.err; echo "align1 exit $?"
timeout 100 $B > $out/bench_c3_sel4_align0_$tag.json 2> $out/bench_c3_sel4_align0_$tag.err; echo "align0 exit $?"
python - <<PY
import json, glob
for f in sorted(glob.glob('$out/bench_c3_*_$tag.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['ms_per_step'], d['kernel_ms'], d['e2e']['ms_per_step'], d['roofline']['kernel'], d['parity_sample']['ok'], d['parity_sample']['max_abs_err'])
    except Exception as e: print('no bench line', f, e)
PY
