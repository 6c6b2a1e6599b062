#!/bin/bash
# round 2, 4-GPU call: 2-rank correctness (allreduce buckets, DDP vs Trainer), NCCL-in-CUDA-graph check, BASELINE configs[4]
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/m4_gpus.txt
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q --tb=short -s > gpurun_out/m4_pytest_dist.log 2>&1
tail -12 gpurun_out/m4_pytest_dist.log
# does the graphed step (NCCL allreduce inside the capture) work on 2 ranks?  bounded: a hang must not take the box down
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 2 --steps 10 --warmup 3 --preheat 4 > gpurun_out/m4_bench_2gpu_graph.json 2> gpurun_out/m4_bench_2gpu_graph.err
echo "rc=$?" >> gpurun_out/m4_bench_2gpu_graph.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus 2 --steps 10 --warmup 3 --preheat 4 --no-graph > gpurun_out/m4_bench_2gpu_eager.json 2> gpurun_out/m4_bench_2gpu_eager.err
timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 \
    tools/run_cfg5.py --iters 200 --every 50 > gpurun_out/cfg5_msvr310_4gpu.json 2> gpurun_out/cfg5_msvr310_4gpu.err
tail -3 gpurun_out/cfg5_msvr310_4gpu.err
ls -la gpurun_out
