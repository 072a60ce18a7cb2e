#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): the C2 bench under torchrun at N ranks, plus N=1 for comparison.
N=${1:-2}
tag=${2:-mg}
out=gpurun_out
mkdir -p $out
set -x
nvidia-smi -L > $out/gpus_$tag.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --forest-cache /tmp/c2.npz > $out/bench_c2_n1_$tag.json 2> $out/bench_c2_n1_$tag.err; echo "n1 exit $?"; tail -1 $out/bench_c2_n1_$tag.json | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 --forest-cache /tmp/c2.npz > $out/bench_c2_n${N}_$tag.json 2> $out/bench_c2_n${N}_$tag.err; echo "n$N exit $?"; tail -1 $out/bench_c2_n${N}_$tag.json | cut -c1-200; tail -5 $out/bench_c2_n${N}_$tag.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > $out/bench_ref_n${N}_$tag.json 2> $out/bench_ref_n${N}_$tag.err; echo "ref n$N exit $?"; tail -1 $out/bench_ref_n${N}_$tag.json | cut -c1-200
