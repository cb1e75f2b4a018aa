#!/bin/bash
# Multi-GPU check on one box: NCCL tests + bench at N ranks (both reduce modes) + N=1 for the ratio.
# usage (under gpurun --gpus N): tools/gpu_scale.sh <tag> <N>
tag=${1:-scale}
n=${2:-2}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -x -q > $out/pytest_nccl.log 2>&1
echo "pytest exit $?" >> $out/pytest_nccl.log
tail -4 $out/pytest_nccl.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err
cat $out/bench_n1.json
for mode in overlap post; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 30 --warmup 5 --reduce $mode > $out/bench_n${n}_$mode.json 2> $out/bench_n${n}_$mode.err
  grep '^{' $out/bench_n${n}_$mode.json; tail -3 $out/bench_n${n}_$mode.err
done
