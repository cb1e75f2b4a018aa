#!/bin/bash
# lean 8-GPU point: N=1, p2p, p2p-overlap (8 CTAs). usage: tools/gpu_p2po8.sh <tag> <N>
tag=${1:-po8}; n=${2:-8}
out=gpurun_out/$tag
mkdir -p $out
timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err
python tools/benchline.py $out/bench_n1.json
run() {  # name, extra env, mode
  env $2 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 30 --warmup 5 --reduce $3 --no-cpu-baseline > $out/bench_n${n}_$1.json 2> $out/bench_n${n}_$1.err
  echo "exit $? $1"; python tools/benchline.py $out/bench_n${n}_$1.json; grep -i "error\|unavailable" $out/bench_n${n}_$1.err | head -3 | cut -c1-300
}
run p2po8 SVGIR_PEER_BG_GRID=8 p2p-overlap
run p2p X=1 p2p
