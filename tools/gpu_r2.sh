#!/bin/bash
# Round-2 GPU call: GPU parity tests, the default bench line (with its extra keys), the reference arm, and the ncu
# launch list + one full capture of the step's dominant kernels. Everything lands in gpurun_out/<tag>/.
# usage: tools/gpu_r2.sh <tag> [pytest -k expr|all] [skip-ncu] [skip-ref]
tag=${1:-r02}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/gpu.txt 2>&1
nproc >> $out/gpu.txt
if [ "$2" != "none" ]; then
  if [ -n "$2" ] && [ "$2" != "all" ]; then
    timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -k "$2" --durations=15 > $out/pytest_gpu.log 2>&1
  else
    timeout 1700 python -m pytest tests -m gpu -q --maxfail=30 --durations=15 > $out/pytest_gpu.log 2>&1
  fi
  echo "pytest exit $?" >> $out/pytest_gpu.log
  tail -40 $out/pytest_gpu.log
fi
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_ours.json 2> $out/bench_ours.err
echo "bench exit $?"; cat $out/bench_ours.json; tail -5 $out/bench_ours.err
if [ "$4" != "skip-ref" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
  cat $out/bench_ref.json; tail -3 $out/bench_ref.err
fi
if [ "$3" != "skip-ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $out/ncu_launch.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'composite|shade_' -c 8 \
      -o $out/step_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $out/ncu_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bvh_trace_opacity' -c 1 \
      -o $out/bvh_trace_full -f python bench.py --workload visibility > $out/ncu_bvh.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'radiance_(cache|select|loss_fwd)' -c 4 \
      -o $out/radiance_full -f python bench.py --workload radiance --steps 1 --warmup 1 > $out/ncu_radiance.log 2>&1
  ls -la $out
fi
