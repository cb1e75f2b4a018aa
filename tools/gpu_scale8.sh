#!/bin/bash
# N-GPU scaling point: N=1 line for the ratio, then N ranks per reduce mode. usage: tools/gpu_scale8.sh <tag> <N> [mode[:bg_ctas] ...]
tag=${1:-s8}; n=${2:-8}; shift; shift
modes=${@:-post overlap:4}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err
python tools/benchline.py $out/bench_n1.json
for m in $modes; do
  mode=${m%%:*}; ctas=4; [ "$m" != "$mode" ] && ctas=${m##*:}
  NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 30 --warmup 5 --reduce $mode --bg-ctas $ctas > $out/bench_n${n}_${mode}_$ctas.json 2> $out/bench_n${n}_${mode}_$ctas.err
  echo "exit $? $m"; python tools/benchline.py $out/bench_n${n}_${mode}_$ctas.json; grep -i "error\|NCCL WARN" $out/bench_n${n}_${mode}_$ctas.err | head -3 | cut -c1-300
done
