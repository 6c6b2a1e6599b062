#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/graph_probe.py > gpurun_out/j_graph_probe.txt 2>&1
