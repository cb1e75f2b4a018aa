#!/bin/bash
# Quick GPU check: selected parity tests + bench (fused / torch tail). usage: tools/gpu_check.sh <tag> [pytest -k expr]
tag=${1:-chk}
out=gpurun_out/$tag
mkdir -p $out
[ -x tools/bin/ffma2_bench ] && tools/bin/ffma2_bench > $out/ffma2.txt 2>&1 && cat $out/ffma2.txt
timeout 1200 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > $out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/pytest_gpu.log
tail -15 $out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_ours.json 2> $out/bench_ours.err
cat $out/bench_ours.json; tail -3 $out/bench_ours.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --torch-loss > $out/bench_torchloss.json 2> $out/bench_torchloss.err
cat $out/bench_torchloss.json
