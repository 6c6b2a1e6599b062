#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/step_trace.py gpurun_out/step_trace.txt > gpurun_out/step_trace.log 2>&1
tail -n 5 gpurun_out/step_trace.log; tail -n 3 gpurun_out/step_trace.txt
