#!/bin/bash
# Round 2, final evidence at HEAD: smoke, whole GPU suite, the driver-shaped bench (both arms), ncu launch list
# and full captures of one C3 step in both modes.  gpurun brings back at most 64 MiB: the captures are
# exported to CSV / summaries on the box and the .ncu-rep files are dropped when they are too large.
tag=${1:-r2_final}
fastk=${2:-"apply_kernel|quantile_select2"}
out=gpurun_out
mkdir -p $out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/host_$tag.txt; nproc >> $out/host_$tag.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke exit $?"; tail -2 $out/smoke_$tag.log
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest_gpu_$tag.log
timeout 1200 python bench.py --impl reference > $out/bench_ref_c3_$tag.json 2> $out/bench_ref_c3_$tag.err; echo "ref exit $?"; tail -1 $out/bench_ref_c3_$tag.json | cut -c1-300
timeout 1500 python bench.py > $out/bench_c3_$tag.json 2> $out/bench_c3_$tag.err; echo "bench exit $?"; tail -3 $out/bench_c3_$tag.err; tail -1 $out/bench_c3_$tag.json | cut -c1-400
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-estimator-e2e --single-mode"
timeout 900 $B --mode fast > $out/plain_fast_$tag.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c3_fast_$tag.csv $B --mode fast > $out/ncu_launch_$tag.log 2>&1; echo "ncu launches exit $?"
timeout 1500 ncu --set full --clock-control none -k regex:"$fastk" -s 8 -c 8 -f -o $out/prof_c3_fast_step_$tag $B --mode fast > $out/ncu_fast_$tag.log 2>&1; echo "ncu fast exit $?"
timeout 900 $B --mode exact > $out/plain_exact_$tag.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none -k regex:"quantile_bucket" -s 4 -c 4 -f -o $out/prof_c3_exact_step_$tag $B --mode exact > $out/ncu_exact_$tag.log 2>&1; echo "ncu exact exit $?"
for m in fast exact; do
  python tools/ncu_summary.py $out/prof_c3_${m}_step_$tag.ncu-rep > $out/ncu_full_c3_${m}_step_$tag.txt 2>&1
done
total=$(du -sm $out | cut -f1)
if [ "$total" -gt 55 ]; then rm -f $out/prof_c3_exact_step_$tag.ncu-rep; fi
total=$(du -sm $out | cut -f1)
if [ "$total" -gt 55 ]; then rm -f $out/prof_c3_fast_step_$tag.ncu-rep; fi
ls -la $out/*$tag*
du -sh $out
