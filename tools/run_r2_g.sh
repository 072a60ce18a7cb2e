#!/bin/bash
# Round 2, pass G: whole GPU suite; C3 bench (default fast kernel = select2) and select3 in both CTA shapes; streaming.
tag=${1:-r2_g}
out=gpurun_out
mkdir -p $out
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -15 $out/pytest_gpu_$tag.log
timeout 1500 python bench.py --no-cpu-baseline --single-mode > $out/bench_c3_$tag.json 2> $out/bench_c3_$tag.err; echo "bench exit $?"; tail -3 $out/bench_c3_$tag.err
for v in 1 0; do
timeout 900 python bench.py --no-cpu-baseline --single-mode --no-estimator-e2e --option select_impl=3 --option select3_variant=$v > $out/bench_c3_sel3v${v}_$tag.json 2> $out/bench_c3_sel3v${v}_$tag.err; echo "sel3 v$v exit $?"
done
python - <<PY
import json
for f in ('bench_c3_$tag.json','bench_c3_sel3v1_$tag.json','bench_c3_sel3v0_$tag.json'):
    try:
        d=json.loads(open('$out/'+f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['kernel_ms'], d['e2e']['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['parity_sample']['ok'], d.get('e2e_estimator'))
    except Exception as e: print('no bench line', f, e)
PY
timeout 900 python tools/bench_stream.py --config c3 --devices 0 --rows 4000000 > $out/stream_c3_1gpu_$tag.json 2> $out/stream_c3_1gpu_$tag.err; echo "stream1 exit $?"; cat $out/stream_c3_1gpu_$tag.json | cut -c1-700
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-estimator-e2e --single-mode --option select_impl=3 --option select3_variant=1"
timeout 900 $B > $out/plain_sel3v1_$tag.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"quantile_select3" -s 4 -c 1 -f -o $out/prof_c3_sel3v1_$tag $B > $out/ncu_sel3v1_$tag.log 2>&1; echo "ncu exit $?"
ls -la $out/*$tag*
