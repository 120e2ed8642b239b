#!/bin/bash
# Second re-validation call: GPU tests, default bench line (+ 40-step rollout), AuroraWave workload.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider ) > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu2.log
( time timeout 600 python bench.py --rollout 40 ) > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/bench2.json; tail -3 gpurun_out/bench2.err
( time timeout 600 python bench.py --workload aurora-wave-0.25deg-721x1440x13L --steps 5 --no-cpu-baseline ) > gpurun_out/bench_wave.json 2> gpurun_out/bench_wave.err; echo "wave rc=$?"
cut -c1-300 gpurun_out/bench_wave.json; tail -3 gpurun_out/bench_wave.err
