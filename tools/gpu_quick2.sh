#!/bin/bash
# Quick round-2 iteration: selected GPU tests + the headline bench without the extra workloads.
# usage: tools/gpu_quick2.sh <tag> [pytest -k expr] [ncu regex]
tag=${1:-q}
out=gpurun_out/$tag
mkdir -p $out
if [ -n "$2" ] && [ "$2" != "none" ]; then
  timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "$2" > $out/pytest.log 2>&1
  echo "pytest exit $?" >> $out/pytest.log
  tail -15 $out/pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $out/bench.json 2> $out/bench.err
python - <<PY
import json
try:
    d = json.load(open("$out/bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print(d["kernels_ms"])
except Exception as e:
    print("bench failed", e); print(open("$out/bench.err").read()[-1500:])
PY
if [ -n "$3" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$3" -c 4 \
      -o $out/prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $out/ncu.log 2>&1
  ls -la $out
fi
