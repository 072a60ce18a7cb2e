#!/bin/bash
# Full-size C3 (ETQR T=500 msl=5, 1M x 50 train, 1M test rows): bench lines in both modes, ncu of one chunk.
tag=${1:-c3}
out=gpurun_out
mkdir -p $out
set -x
timeout 2400 python bench.py --config c3 --steps 5 --warmup 3 --mode fast --forest-cache /tmp/c3.npz --cpu-seconds 20 > $out/bench_c3_fast_$tag.json 2> $out/bench_c3_fast_$tag.err; echo "fast exit $?"; tail -2 $out/bench_c3_fast_$tag.err; tail -1 $out/bench_c3_fast_$tag.json | cut -c1-400
timeout 900 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline --mode exact --forest-cache /tmp/c3.npz > $out/bench_c3_exact_$tag.json 2> $out/bench_c3_exact_$tag.err; echo "exact exit $?"; tail -1 $out/bench_c3_exact_$tag.json | cut -c1-400
timeout 900 ncu --set full --clock-control none -k regex:"apply_kernel|quantile_select2|quantile_bucket" -s 3 -c 3 -f -o $out/prof_c3_$tag \
    python tools/sweep.py --config c3 --mode fast --reps 1 --rows 262144 --grid '[{}]' --forest-cache /tmp/c3.npz > $out/ncu_c3_$tag.log 2>&1; echo "ncu exit $?"
