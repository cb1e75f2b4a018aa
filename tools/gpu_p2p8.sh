#!/bin/bash
# 8-GPU point: peer all-reduce test (2 ranks), all-reduce microbenchmark, bench N=1 / N post / N p2p. usage: tools/gpu_p2p8.sh <tag> <N>
tag=${1:-p8}; n=${2:-8}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -x -q -k peer > $out/pytest_peer.log 2>&1
echo "pytest exit $?" >> $out/pytest_peer.log; tail -3 $out/pytest_peer.log | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 tools/peer_bench.py 2> $out/peer_bench.err | tee $out/peer_bench.json | cut -c1-1500
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err
python tools/benchline.py $out/bench_n1.json
run() {  # name, extra env, mode
  env $2 NCCL_DEBUG=WARN timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 30 --warmup 5 --reduce $3 > $out/bench_n${n}_$1.json 2> $out/bench_n${n}_$1.err
  echo "exit $? $1"; python tools/benchline.py $out/bench_n${n}_$1.json; grep -i "error\|NCCL WARN\|unavailable" $out/bench_n${n}_$1.err | head -3 | cut -c1-300
}
run p2p_mc SVGIR_PEER_MULTICAST=1 p2p
run post X=1 post
