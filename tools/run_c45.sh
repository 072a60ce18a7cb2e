#!/bin/bash
# Reduced C4 (99 quantiles per row) and C5 (T=1000, 64 features, rows through the host pipeline) bench lines.
tag=${1:-c45}
out=gpurun_out
mkdir -p $out
set -x
timeout 1200 python bench.py --config c4q --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_c4q_$tag.json 2> $out/bench_c4q_$tag.err; echo "c4q exit $?"; tail -2 $out/bench_c4q_$tag.err; tail -1 $out/bench_c4q_$tag.json | cut -c1-300
timeout 1500 python bench.py --config c5q --steps 3 --warmup 3 --mode fast --no-cpu-baseline > $out/bench_c5q_fast_$tag.json 2> $out/bench_c5q_fast_$tag.err; echo "c5q exit $?"; tail -2 $out/bench_c5q_fast_$tag.err; tail -1 $out/bench_c5q_fast_$tag.json | cut -c1-300
