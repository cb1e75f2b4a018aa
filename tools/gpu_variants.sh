#!/bin/bash
# Runs the headline bench (no extras) once per variant: "name:LIB:ENVVAR=VALUE" triples, LIB = variants/<x>.so or "-"
# usage: tools/gpu_variants.sh <tag> name:lib:env ...
tag=$1; shift
out=gpurun_out/$tag
mkdir -p $out
for spec in "$@"; do
  IFS=: read name lib envs <<< "$spec"
  libpath=""
  [ "$lib" != "-" ] && libpath="$PWD/svg-ir_b200/svgir_b200/variants/$lib.so"
  env SVGIR_B200_LIB="$libpath" $envs timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-e2e > $out/$name.json 2> $out/$name.err
  python - <<PY
import json
try:
    d = json.load(open("$out/$name.json"))
    k = d["kernels_ms"]
    print("%-14s %7.1f it/s  %.4f ms | bwd %.4f fwd %.4f shf %.4f shb %.4f" % ("$name", d["value"], d["ms_per_step"], k["composite_bwd"], k["composite_fwd"], k["shade_fwd"], k["shade_bwd"]))
except Exception as e:
    print("$name failed", e); print(open("$out/$name.err").read()[-800:])
PY
done
