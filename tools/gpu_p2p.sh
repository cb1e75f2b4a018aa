#!/bin/bash
# peer all-reduce check at N ranks: correctness test, then bench per mode. usage: tools/gpu_p2p.sh <tag> <N> [skip-test]
tag=${1:-p2p}; n=${2:-2}
out=gpurun_out/$tag
mkdir -p $out
if [ "$3" != "skip-test" ]; then
  timeout 420 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -x -q -k peer > $out/pytest_peer.log 2>&1
  echo "pytest exit $?" >> $out/pytest_peer.log; tail -25 $out/pytest_peer.log | cut -c1-400
fi
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err
python tools/benchline.py $out/bench_n1.json
run() {  # name, extra env, mode
  env $2 NCCL_DEBUG=WARN timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 30 --warmup 5 --reduce $3 > $out/bench_n${n}_$1.json 2> $out/bench_n${n}_$1.err
  echo "exit $? $1"; python tools/benchline.py $out/bench_n${n}_$1.json; grep -i "error\|NCCL WARN\|unavailable" $out/bench_n${n}_$1.err | head -3 | cut -c1-300
}
run post X=1 post
run p2p_mc SVGIR_PEER_MULTICAST=1 p2p
run p2p_ldst SVGIR_PEER_MULTICAST=0 p2p
