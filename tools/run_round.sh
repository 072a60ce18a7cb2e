#!/bin/bash
# One GPU-box pass: parity tests, C2 bench (both arms), ncu launch list and full captures.
# Usage (from the repo root, under gpurun): bash tools/run_round.sh <tag>
tag=${1:-vX}
out=gpurun_out
mkdir -p $out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/host_$tag.txt; nproc >> $out/host_$tag.txt
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke exit $?"
python bench.py --forest-cache /tmp/c2.npz > $out/bench_c2_$tag.json 2> $out/bench_c2_$tag.err; echo "bench exit $?"; tail -1 $out/bench_c2_$tag.json | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 --forest-cache /tmp/c2.npz > $out/bench_ref_c2_$tag.json 2> $out/bench_ref_c2_$tag.err; echo "ref exit $?"
# launch list (cold, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c2_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --forest-cache /tmp/c2.npz > $out/ncu_launch_$tag.log 2>&1; echo "ncu launches exit $?"
# full capture of the two dominant kernels of the same command
ncu --set full --clock-control none --import-source on -k regex:"apply_kernel|quantile_warp" -s 4 -c 2 -f -o $out/prof_c2_$tag \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --forest-cache /tmp/c2.npz > $out/ncu_full_$tag.log 2>&1; echo "ncu full exit $?"
