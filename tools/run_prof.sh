set -x
python tools/sweep.py --config c3q --rows 16384 --reps 1 --mode fast --grid "[{}]" --forest-cache /tmp/c3q.npz > gpurun_out/sweep_c3q_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:quantile_select -s 2 -c 1 -o gpurun_out/prof_select_c3q python tools/sweep.py --config c3q --rows 16384 --reps 1 --mode fast --grid "[{}]" --forest-cache /tmp/c3q.npz > gpurun_out/ncu_c3q.log 2>&1
tail -3 gpurun_out/sweep_c3q_small.log; tail -3 gpurun_out/ncu_c3q.log
