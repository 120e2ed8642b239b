#!/bin/bash
# One gpurun call that re-validates the tree on a B200: smoke, GPU parity tests, the bench line and the ncu
# metrics pass (launch list + DRAM traffic).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json | cut -c1-600
( time timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'gemm|window_attention|ln_mod|patch|perceiver|linear_small' --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e ) > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/launches.csv
tail -3 gpurun_out/ncu_bench.log | cut -c1-300
