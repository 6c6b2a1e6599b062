#!/bin/bash
# round 2, 8-GPU call: BASELINE configs[2] (RGBNT100 EDITOR.yml, B=128 per rank, bf16, 8 x B200) and configs[1] at N=8
set -x
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 8 --config RGBNT100 --steps 20 --warmup 5 > gpurun_out/m8_bench_rgbnt100_8gpu.json 2> gpurun_out/m8_bench_rgbnt100_8gpu.err
echo "rc=$?" >> gpurun_out/m8_bench_rgbnt100_8gpu.err
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus 8 --config RGBNT201 --steps 20 --warmup 5 > gpurun_out/m8_bench_rgbnt201_8gpu.json 2> gpurun_out/m8_bench_rgbnt201_8gpu.err
echo "rc=$?" >> gpurun_out/m8_bench_rgbnt201_8gpu.err
ls -la gpurun_out
