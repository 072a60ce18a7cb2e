#!/bin/bash
# Round 2, last pass: config 2 bench line at HEAD (continuity with round 1) and one more fuzz seed.
tag=${1:-r2_last}
out=gpurun_out
mkdir -p $out
set -x
timeout 300 python bench.py --config c2 > $out/bench_c2_$tag.json 2> $out/bench_c2_$tag.err; echo "c2 exit $?"; tail -1 $out/bench_c2_$tag.json | cut -c1-300
timeout 200 python tools/fuzz_parity.py --cases 30 --seed 12 > $out/fuzz_seed12_$tag.log 2>&1; echo "fuzz exit $?"; tail -2 $out/fuzz_seed12_$tag.log | cut -c1-300
