#!/bin/bash
# Round profiling recipe (run under gpurun, 1 GPU): tests, bench (own + reference arm), ncu launch list of one step, ncu
# --set full captures of the dominant kernels, in-step kernel timeline, SFTS isolation.  tools/summarize_profiles.py turns
# the results into the text files committed under profiles/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 5 > gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_own.json 2> gpurun_out/bench_own.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.txt 2>&1
timeout 300 python tools/attn_bench.py > gpurun_out/attn_bench.txt 2>&1
timeout 600 python tools/step_trace.py gpurun_out/step_trace.txt > gpurun_out/step_trace.log 2>&1
# (tools/one_step.py = 4 training steps and nothing else; the summariser takes the step between the last two SGD kernels)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python tools/one_step.py 4 > gpurun_out/bench_under_ncu.log 2>&1
# the GEMM in the step: six consecutive launches of the third training step (fc2 dgrad, wgrads, fc1 dgrad ...)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 420 -c 6 -f -o gpurun_out/prof_gemm \
    python tools/one_step.py 3 > /dev/null 2>&1
# single shapes in isolation (cold L2): fc1 forward (GELU + GELU'), fc2 dgrad (x saved GELU'), proj forward (fp32 residual),
# fc1 dgrad (plain bf16 store, mainloop-bound)
for w in fc1 fc2d proj fc1d; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 2 -c 1 -f -o gpurun_out/prof_gemm_$w \
     python tools/gemm_one.py $w > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -s 2 -c 1 -f -o gpurun_out/prof_attn_fwd \
    python tools/attn_bench.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd -s 2 -c 1 -f -o gpurun_out/prof_attn_bwd \
    python tools/attn_bench.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rollout|freq_counts|ln_bwd_kernel|ln_fwd_kernel" -s 60 -c 4 -f -o gpurun_out/prof_sfts_ln \
    python tools/one_step.py 2 > /dev/null 2>&1
timeout 600 python tools/sfts_bench.py > gpurun_out/sfts_bench.json 2> gpurun_out/sfts_bench.err
ls -la gpurun_out
