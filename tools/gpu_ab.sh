#!/bin/bash
# A/B of two builds of the library: parity tests on the in-tree build, then the bench on each.
# usage: tools/gpu_ab.sh <tag> [alt-lib.so ...]
tag=${1:-ab}; shift
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1
echo "pytest exit $?" >> $out/pytest.log
tail -4 $out/pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $out/bench_A.json 2> $out/bench_A.err
python - $out/bench_A.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("A", d["value"], d["ms_per_step"], d["kernels_ms"])
PY
for lib in "$@"; do
  n=$(basename $lib .so)
  SVGIR_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $out/bench_$n.json 2> $out/bench_$n.err
  python - $out/bench_$n.json $n <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d["value"], d["ms_per_step"], d["kernels_ms"])
PY
done
