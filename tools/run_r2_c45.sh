#!/bin/bash
# Round 2: BASELINE.json configs[3] (C4 at 1M rows + the depth-10 stress variant) and configs[4] (C5: 10M streamed rows) on one GPU.
tag=${1:-r2_c45}
out=gpurun_out
mkdir -p $out
set -x
timeout 2400 python bench.py --config c5 --steps 3 --warmup 3 --cpu-seconds 6 > $out/bench_c5_$tag.json 2> $out/bench_c5_$tag.err; echo "c5 exit $?"; tail -3 $out/bench_c5_$tag.err; tail -1 $out/bench_c5_$tag.json | cut -c1-400
timeout 900 python tools/bench_stream.py --config c5 --devices 0 --rows 10000000 --chunk-rows 1000000 --reps 1 > $out/stream_c5_1gpu_$tag.json 2> $out/stream_c5_1gpu_$tag.err; echo "stream c5 exit $?"; cat $out/stream_c5_1gpu_$tag.json | cut -c1-600
timeout 2400 python bench.py --config c4 --mode exact --single-mode --steps 5 --warmup 3 --cpu-seconds 6 > $out/bench_c4_$tag.json 2> $out/bench_c4_$tag.err; echo "c4 exit $?"; tail -3 $out/bench_c4_$tag.err; tail -1 $out/bench_c4_$tag.json | cut -c1-400
timeout 1200 python bench.py --config c4_depth10 --mode exact --single-mode --steps 3 --warmup 3 --no-cpu-baseline --no-estimator-e2e > $out/bench_c4d10_$tag.json 2> $out/bench_c4d10_$tag.err; echo "c4d10 exit $?"; tail -3 $out/bench_c4d10_$tag.err; tail -1 $out/bench_c4d10_$tag.json | cut -c1-400
ls -la $out/*$tag*
