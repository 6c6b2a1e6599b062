#!/bin/bash
# Round profiling recipe (run under gpurun, 1 GPU): tests, bench (own + reference arm), ncu launch list of one step, ncu
# --set full captures of the dominant kernels, in-step kernel timeline.  tools/summarize_profiles.py <tag> turns the results
# into the text files committed under profiles/.
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rs 2>&1 | tail -n 12 > gpurun_out/pytest_gpu.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_own.json 2> gpurun_out/bench_own.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python bench.py --config RGBNT100 --steps 20 --warmup 5 --no-cpu-baseline --no-sfts > gpurun_out/bench_rgbnt100.json 2> gpurun_out/bench_rgbnt100.err
timeout 900 python bench.py --config MSVR310 --precision fp32 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_msvr310_fp32.json 2> gpurun_out/bench_msvr310_fp32.err
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.txt 2>&1
timeout 300 python tools/attn_bench.py > gpurun_out/attn_bench.txt 2>&1
timeout 300 python tools/ln_bench.py > gpurun_out/ln_bench.txt 2>&1
timeout 600 python tools/step_trace.py gpurun_out/step_trace.txt > gpurun_out/step_trace.log 2>&1
# (tools/one_step.py = 4 eager training steps and nothing else; the summariser takes the step between the last two SGD kernels)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python tools/one_step.py 4 > gpurun_out/bench_under_ncu.log 2>&1
# the GEMM in the step: six consecutive launches of the third training step (fc2 dgrad, wgrads, fc1 dgrad ...)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 420 -c 6 -f -o gpurun_out/prof_gemm \
    python tools/one_step.py 3 > /dev/null 2>&1
for w in fc1 fc2d proj fc1d; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 2 -c 1 -f -o gpurun_out/prof_gemm_$w \
     python tools/gemm_one.py $w > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -s 2 -c 1 -f -o gpurun_out/prof_attn_fwd \
    python tools/attn_bench.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd -s 2 -c 1 -f -o gpurun_out/prof_attn_bwd \
    python tools/attn_bench.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ln_bwd_kernel|ln_fwd_kernel" -s 11 -c 2 -f -o gpurun_out/prof_ln \
    python tools/ln_bench.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rollout_topk|freq_counts|sfts_pack_fwd" -s 3 -c 3 -f -o gpurun_out/prof_sfts \
    python tools/sfts_bench.py --batch 256 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"aug_main_kernel|aug_hpass_kernel" -c 2 -f -o gpurun_out/prof_augment \
    python -m pytest tests/test_augment_gpu.py -q -k "300-140" > /dev/null 2>&1
ls -la gpurun_out
