#!/bin/bash
# One gpurun call: GPU parity tests, the three bench arms, the ncu launch list of the bench command and
# one full ncu capture of the step's kernels. Everything lands in gpurun_out/<tag>/.
# usage: tools/gpu_round.sh <tag> [skip-ncu]
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/gpu.txt 2>&1
nproc >> $out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench_ours.json 2> $out/bench_ours.err
cat $out/bench_ours.json
timeout 600 python bench.py --impl reference_cuda --steps 10 --warmup 3 > $out/bench_refcuda.json 2> $out/bench_refcuda.err
cat $out/bench_refcuda.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
cat $out/bench_ref.json
timeout 300 python bench.py --workload relight --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_relight.json 2> $out/bench_relight.err
cat $out/bench_relight.json
if [ "$2" != "skip-ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $out/ncu_launch.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'composite|shade' -c 8 \
      -o $out/step_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/ncu_full.log 2>&1
  ls -la $out
fi
