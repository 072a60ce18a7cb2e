#!/bin/bash
# Round 2, pass A: HEAD kernels under the new bench (C3 default, strong scaling, parity sample),
# GPU tests, and ncu of one step (launch list + full capture of K1 / K2-fast / K2-exact).
tag=${1:-r2_a}
out=gpurun_out
mkdir -p $out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/host_$tag.txt; nproc >> $out/host_$tag.txt; free -g >> $out/host_$tag.txt; df -h /tmp >> $out/host_$tag.txt
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest_gpu_$tag.log
timeout 1500 python bench.py > $out/bench_c3_$tag.json 2> $out/bench_c3_$tag.err; echo "bench exit $?"; tail -5 $out/bench_c3_$tag.err; tail -1 $out/bench_c3_$tag.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_c3_$tag.json 2> $out/bench_ref_c3_$tag.err; echo "ref exit $?"; tail -1 $out/bench_ref_c3_$tag.json | cut -c1-300
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-estimator-e2e --single-mode"
timeout 900 $B --mode fast > $out/plain_fast_$tag.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"apply_kernel|quantile_select2" -s 8 -c 2 -f -o $out/prof_c3_fast_$tag \
    $B --mode fast > $out/ncu_fast_$tag.log 2>&1; echo "ncu fast exit $?"
timeout 900 $B --mode exact > $out/plain_exact_$tag.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"quantile_bucket" -s 4 -c 1 -f -o $out/prof_c3_exact_$tag \
    $B --mode exact > $out/ncu_exact_$tag.log 2>&1; echo "ncu exact exit $?"
ls -la $out/*$tag*
