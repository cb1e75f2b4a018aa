#!/bin/bash
# Quick GPU iteration: selected tests, graph + eager bench lines, step timelines. usage: tools/gpu_quick.sh <tag> [pytest -k expr]
tag=${1:-q}
out=gpurun_out/$tag
mkdir -p $out
if [ -n "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > $out/pytest.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1
fi
echo "pytest exit $?" >> $out/pytest.log
tail -15 $out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_graph.json 2> $out/bench_graph.err
cat $out/bench_graph.json; tail -5 $out/bench_graph.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --eager > $out/bench_eager.json 2> $out/bench_eager.err
cat $out/bench_eager.json; tail -5 $out/bench_eager.err
timeout 300 python tools/timeline.py --graph > $out/timeline_graph.txt 2> $out/timeline_graph.err
tail -25 $out/timeline_graph.txt; tail -5 $out/timeline_graph.err
timeout 300 python tools/timeline.py > $out/timeline_eager.txt 2> $out/timeline_eager.err
tail -4 $out/timeline_eager.txt | head -2
