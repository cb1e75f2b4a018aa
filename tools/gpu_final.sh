#!/bin/bash
# Final single-GPU evidence without the long full captures: GPU tests, smoke, default line, reference arm, ncu launch list.
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/gpu.txt 2>&1; nproc >> $out/gpu.txt
timeout 1700 python -m pytest tests -m gpu -q --maxfail=30 --durations=10 > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log; tail -4 $out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > $out/smoke.log 2>&1; tail -1 $out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_ours.json 2> $out/bench_ours.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $out/ncu_launch.log 2>&1
python - <<PY
import json
d=json.load(open("$out/bench_ours.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["kernels_ms"], d["roofline"])
print({k:(d[k].get("value") if isinstance(d.get(k),dict) else d.get(k)) for k in ("c4","c5","relight","reference_cuda","cpu_baseline","extras_error")})
PY
