#!/bin/bash
# Round 2, multi-GPU pass (gpurun --gpus 2): devices= product API, torchrun strong scaling of C3, streaming over 2 GPUs.
tag=${1:-r2_mg}
out=gpurun_out
mkdir -p $out
set -x
nvidia-smi --query-gpu=index,name --format=csv > $out/host_$tag.txt; nproc >> $out/host_$tag.txt
timeout 600 python -m pytest tests/test_select3.py -x -q -k "devices or two_host_threads" > $out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest_$tag.log
timeout 1500 python bench.py --gpus 1 --no-cpu-baseline --steps 10 > $out/bench_c3_n1_$tag.json 2> $out/bench_c3_n1_$tag.err; echo "n1 exit $?"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --steps 10 > $out/bench_c3_n2_$tag.json 2> $out/bench_c3_n2_$tag.err; echo "n2 exit $?"; tail -3 $out/bench_c3_n2_$tag.err
python - <<PY
import json
for n in (1, 2):
    try:
        d=json.loads(open('$out/bench_c3_n%d_$tag.json' % n).read().strip().splitlines()[-1])
        print(n, d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'], d['parity_sample'] and d['parity_sample']['ok'])
    except Exception as e: print('no bench line', n, e)
PY
timeout 900 python tools/bench_stream.py --config c3 --devices 0 --rows 4000000 > $out/stream_c3_1gpu_$tag.json 2> $out/stream_c3_1gpu_$tag.err; echo "stream1 exit $?"; cat $out/stream_c3_1gpu_$tag.json | cut -c1-600
timeout 900 python tools/bench_stream.py --config c3 --devices 0,1 --rows 4000000 > $out/stream_c3_2gpu_$tag.json 2> $out/stream_c3_2gpu_$tag.err; echo "stream2 exit $?"; cat $out/stream_c3_2gpu_$tag.json | cut -c1-600
ls -la $out/*$tag*
