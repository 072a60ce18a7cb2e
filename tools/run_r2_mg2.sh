#!/bin/bash
# Round 2, final multi-GPU check at HEAD (gpurun --gpus 2): the driver's launch line for N=2, both arms.
tag=${1:-r2_mg2}
out=gpurun_out
mkdir -p $out
set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_c3_n2_$tag.json 2> $out/bench_c3_n2_$tag.err; echo "n2 exit $?"; tail -3 $out/bench_c3_n2_$tag.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $out/bench_ref_c3_n2_$tag.json 2> $out/bench_ref_c3_n2_$tag.err; echo "ref n2 exit $?"; tail -1 $out/bench_ref_c3_n2_$tag.json | cut -c1-200
timeout 600 python tools/bench_stream.py --config c3 --devices 0,1 --rows 4000000 > $out/stream_c3_2gpu_$tag.json 2> $out/stream_c3_2gpu_$tag.err; echo "stream2 exit $?"; cat $out/stream_c3_2gpu_$tag.json | cut -c1-400
python - <<PY
import json
try:
    d=json.loads(open('$out/bench_c3_n2_$tag.json').read().strip().splitlines()[-1])
    print(d['n_gpus'], d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'], d['parity_sample'] and d['parity_sample']['ok'], d.get('cpu_baseline') is not None)
except Exception as e: print('no bench line', e)
PY
ls -la $out/*$tag*
