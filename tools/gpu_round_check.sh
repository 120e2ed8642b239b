#!/bin/bash
# One gpurun call that re-validates the tree on a B200; everything lands in gpurun_out/.  Sections (pick with $1,
# default "smoke tests bench"): smoke | tests | bench | graph | refarm | ncu | ncufull
mkdir -p gpurun_out
SECTIONS="${*:-smoke tests bench}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
for s in $SECTIONS; do
case $s in
smoke)
  ( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
  tail -2 gpurun_out/smoke.log | head -1 ;;
tests)
  ( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider -s -rs ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
  grep -E "passed|failed|FAILED|SKIPPED" gpurun_out/pytest_gpu.log | tail -12
  grep -A14 "full-size parity" gpurun_out/pytest_gpu.log | cut -c1-160 ;;
bench)
  ( time timeout 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  cut -c1-900 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
  python - <<'E'
import json
for ln in open("gpurun_out/bench.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("gpu_reference:", json.dumps(d.get("gpu_reference"))[:700])
        print("cpu_baseline:", json.dumps(d.get("cpu_baseline"))[:700])
E
  ;;
refarm)
  ( time timeout 1200 python bench.py --impl reference --steps 4 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "refarm rc=$?"
  cut -c1-1200 gpurun_out/bench_ref.json ;;
graph)
  ( time timeout 400 python bench.py --steps 10 --warmup 3 --cuda-graph --rollout 40 --no-cpu-baseline --no-gpu-reference ) > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "graph rc=$?"
  python tools/show_bench.py gpurun_out/bench_graph.json ;;
ncufull)
  ( time timeout 600 ncu --set full --clock-control none --import-source on -k regex:'window_attention_tc|gemm2_bf16|gemm_ln|ln_mod_residual' -c 8 -o gpurun_out/r02_prof python tools/ncu_targets.py ) > gpurun_out/ncu_full.log 2>&1; echo "ncufull rc=$?"
  tail -3 gpurun_out/ncu_full.log | cut -c1-300; ls -la gpurun_out/r02_prof.ncu-rep ;;
ncu)
  ( time timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:'gemm|window_attention|ln_mod|patch|perceiver|linear_small|halo' --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-gpu-reference ) > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
  wc -l gpurun_out/launches.csv; tail -3 gpurun_out/ncu_bench.log | cut -c1-300 ;;
esac
done
