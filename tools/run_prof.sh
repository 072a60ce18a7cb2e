set -x
python tools/sweep.py --config c2 --reps 1 --grid "[{}]" --forest-cache /tmp/c2.npz > gpurun_out/sweep_c2_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"apply_kernel|quantile_warp" -s 2 -c 2 -o gpurun_out/prof_c2_v5 python tools/sweep.py --config c2 --reps 1 --grid "[{}]" --forest-cache /tmp/c2.npz > gpurun_out/ncu_c2.log 2>&1
tail -3 gpurun_out/sweep_c2_small.log; tail -3 gpurun_out/ncu_c2.log
