#!/bin/bash
# usage: gpu_multi.sh N   — cfg-4 inference bench and cfg-5 training step under torchrun on N GPUs of one box
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02e_scale_n$N.json 2> gpurun_out/r02e_scale_n$N.err
echo "bench rc=$?"
timeout 600 $TR bench.py --gpus $N --train --verify --cfg cfg5 --steps 10 --warmup 3 > gpurun_out/r02e_train_n$N.json 2> gpurun_out/r02e_train_n$N.err
echo "train rc=$?"
timeout 600 $TR bench.py --gpus $N --cfg cfg3 --steps 10 --warmup 3 > gpurun_out/r02e_weak_n$N.json 2> gpurun_out/r02e_weak_n$N.err
echo "weak rc=$?"
tail -c 1500 gpurun_out/r02e_scale_n$N.json; echo; tail -c 1800 gpurun_out/r02e_train_n$N.json; echo; tail -c 600 gpurun_out/r02e_weak_n$N.json; tail -n 3 gpurun_out/r02e_scale_n$N.err; tail -n 3 gpurun_out/r02e_train_n$N.err; exit 0
