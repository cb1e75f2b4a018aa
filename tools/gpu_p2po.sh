#!/bin/bash
# segment-overlapped peer all-reduce at N ranks: correctness test (2 ranks), then bench p2p vs p2p-overlap. usage: tools/gpu_p2po.sh <tag> <N> [skip-test]
tag=${1:-p2po}; n=${2:-2}
out=gpurun_out/$tag
mkdir -p $out
if [ "$3" != "skip-test" ]; then
  timeout 300 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -x -q -k peer > $out/pytest_peer.log 2>&1
  echo "pytest exit $?" >> $out/pytest_peer.log; tail -25 $out/pytest_peer.log | cut -c1-400
fi
run() {  # name, extra env, mode
  env $2 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 30 --warmup 5 --reduce $3 --no-cpu-baseline > $out/bench_n${n}_$1.json 2> $out/bench_n${n}_$1.err
  echo "exit $? $1"; python tools/benchline.py $out/bench_n${n}_$1.json; grep -i "error\|unavailable" $out/bench_n${n}_$1.err | head -3 | cut -c1-300
}
run p2p X=1 p2p
run p2po16 SVGIR_PEER_BG_GRID=16 p2p-overlap
run p2po8 SVGIR_PEER_BG_GRID=8 p2p-overlap
