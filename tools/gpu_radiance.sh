#!/bin/bash
# usage: tools/gpu_radiance.sh <tag> [pytest -k expr|none] [ncu regex]
tag=${1:-rad}
out=gpurun_out/$tag
mkdir -p $out
if [ "${2:-radiance or optim}" != "none" ]; then
  timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "${2:-radiance or optim}" > $out/pytest.log 2>&1
  echo "pytest exit $?" >> $out/pytest.log
  tail -25 $out/pytest.log
fi
timeout 300 python tools/time_radiance.py > $out/time.json 2> $out/time.err; cat $out/time.json; tail -5 $out/time.err
if [ -n "$3" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$3" -c 4 \
      -o $out/prof -f python tools/time_radiance.py > $out/ncu.log 2>&1
  ls -la $out
fi
