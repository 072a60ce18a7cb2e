set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2_v5.json 2> gpurun_out/bench_c2_v5.err; tail -1 gpurun_out/bench_c2_v5.json | cut -c1-400
timeout 1500 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline --mode fast --forest-cache /tmp/c3.npz > gpurun_out/bench_c3_fast.json 2> gpurun_out/bench_c3_fast.err; tail -2 gpurun_out/bench_c3_fast.err; tail -1 gpurun_out/bench_c3_fast.json
timeout 900 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline --mode exact --forest-cache /tmp/c3.npz > gpurun_out/bench_c3_exact.json 2> gpurun_out/bench_c3_exact.err; tail -2 gpurun_out/bench_c3_exact.err; tail -1 gpurun_out/bench_c3_exact.json
