#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -s 2 -c 1 -f -o gpurun_out/ncu_attn_fwd python tools/attn_bench.py > gpurun_out/ncu_attn_fwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd -s 2 -c 1 -f -o gpurun_out/ncu_attn_bwd python tools/attn_bench.py > gpurun_out/ncu_attn_bwd.log 2>&1
tail -n 3 gpurun_out/ncu_attn_fwd.log gpurun_out/ncu_attn_bwd.log
