#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/graph_probe.py > gpurun_out/k_graph_probe.txt 2>&1
grep "STAGE" gpurun_out/k_graph_probe.txt
timeout 900 python -m pytest tests/test_dropin_do_train_gpu.py -m gpu -q --tb=short -k "graphed or gradscaler or accumulates" > gpurun_out/k_pytest_sel.log 2>&1
tail -5 gpurun_out/k_pytest_sel.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-sfts --no-cpu-baseline > gpurun_out/k_bench_graph.json 2> gpurun_out/k_bench_graph.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-sfts --no-cpu-baseline --no-graph > gpurun_out/k_bench_eager.json 2> gpurun_out/k_bench_eager.err
ls -la gpurun_out
