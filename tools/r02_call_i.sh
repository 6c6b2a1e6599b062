#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/graph_probe.py > gpurun_out/i_graph_probe.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29517 tools/run_cfg5.py --iters 4 --every 2 --distinct-batches 2 > gpurun_out/i_cfg5_smoke.json 2> gpurun_out/i_cfg5_smoke.err
tail -5 gpurun_out/i_cfg5_smoke.err
ls -la gpurun_out
