#!/bin/bash
# 2-GPU call: the 2-rank GPU tests + the headline bench at N=2 (torchrun) and N=1 on the same box.
tag=${1:-n2}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L > $out/gpus.txt
timeout 900 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -q > $out/pytest.log 2>&1
echo "pytest exit $?" >> $out/pytest.log; tail -25 $out/pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $out/bench_n1.json 2> $out/bench_n1.err
N=${2:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $out/bench_n$N.json 2> $out/bench_n$N.err
python - <<PY
import json
for n in (1, $N):
    try:
        t = open("$out/bench_n%d.json" % n).read()
        d = json.loads(t[t.index('{"metric'):])
        print("N=%d value %.1f ms %.4f e2e %s allreduce %s" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"], json.dumps(d.get("grad_allreduce"))[:400]))
        if "c4" in d: print("  c4", json.dumps(d["c4"])[:500])
    except Exception as e:
        print("N=%d failed" % n, e); print(open("$out/bench_n%d.err" % n).read()[-2500:])
PY
