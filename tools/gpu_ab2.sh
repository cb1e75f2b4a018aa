#!/bin/bash
# A/B of the backward compositors: warp-independent (default) vs CTA-staged lane kernel (SVGIR_BWD_CTA=1).
tag=${1:-ab}
out=gpurun_out/$tag
mkdir -p $out
if [ -n "$2" ] && [ "$2" != "none" ]; then
  timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "$2" > $out/pytest.log 2>&1
  echo "pytest exit $?" >> $out/pytest.log
  tail -4 $out/pytest.log
fi
for v in 0 1; do
  SVGIR_BWD_CTA=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-e2e > $out/bench_$v.json 2> $out/bench_$v.err
  python - <<PY
import json
try:
    d = json.load(open("$out/bench_$v.json"))
    print("BWD_CTA=$v value", d["value"], "ms", d["ms_per_step"]); print(d["kernels_ms"])
except Exception as e:
    print("bench failed", e); print(open("$out/bench_$v.err").read()[-1500:])
PY
done
if [ -n "$3" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$3" -c 2 \
      -o $out/prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $out/ncu.log 2>&1
fi
