#!/bin/bash
# Final re-validation call of the round: GPU tests, the bench line, then ncu full captures (attention, GEMM) to mine offline.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider ) > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu3.log | head -2
( time timeout 300 python bench.py ) > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"
cut -c1-260 gpurun_out/bench3.json; tail -2 gpurun_out/bench3.err
timeout 200 ncu --set full --import-source on --clock-control none -k regex:window_attention_tc --launch-skip 3 --launch-count 1 \
    -o gpurun_out/attn_s1_unshifted python tools/attn_probe.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:gemm2 --launch-count 4 \
    -o gpurun_out/gemm_shapes python tools/gemm_once.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep
