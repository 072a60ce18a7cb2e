#!/bin/bash
# Round 2, pass H: rows per kernel chunk (locality of the row order -> L2 hit rate of the leaf lists).
tag=${1:-r2_h}
out=gpurun_out
mkdir -p $out
set -x
B="python bench.py --no-cpu-baseline --single-mode --no-estimator-e2e --steps 10"
for c in 262144 524288 1048576; do
timeout 900 $B --option rows_per_chunk=$c > $out/bench_c3_sel2_chunk${c}_$tag.json 2> $out/bench_c3_sel2_chunk${c}_$tag.err; echo "sel2 chunk $c exit $?"
done
for c in 524288 1048576; do
timeout 900 $B --option rows_per_chunk=$c --option select_impl=3 > $out/bench_c3_sel3_chunk${c}_$tag.json 2> $out/bench_c3_sel3_chunk${c}_$tag.err; echo "sel3 chunk $c exit $?"
done
timeout 900 $B --mode exact --option rows_per_chunk=1048576 > $out/bench_c3_exact_chunk1048576_$tag.json 2> $out/bench_c3_exact_chunk1048576_$tag.err; echo "exact exit $?"
python - <<PY
import json, glob
for f in sorted(glob.glob('$out/bench_c3_*_$tag.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['ms_per_step'], d['kernel_ms'], d['e2e']['ms_per_step'], d['roofline']['kernel'], d['parity_sample']['ok'])
    except Exception as e: print('no bench line', f, e)
PY
B2="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-estimator-e2e --single-mode --option rows_per_chunk=1048576"
timeout 900 $B2 > $out/plain_$tag.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"quantile_select2|apply_kernel" -s 2 -c 2 -f -o $out/prof_c3_chunk1m_$tag $B2 > $out/ncu_$tag.log 2>&1; echo "ncu exit $?"
ls -la $out/*$tag*
