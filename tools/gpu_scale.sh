#!/bin/bash
# N-GPU call without the tests: the default bench line at N (torchrun, all extra keys) and the headline at N=1 on the same box.
# usage: tools/gpu_scale.sh <tag> <N>
tag=${1:-scale}; N=${2:-8}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L > $out/gpus.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $out/bench_n1.json 2> $out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > $out/bench_n$N.json 2> $out/bench_n$N.err
python - <<PY
import json
for n in (1, $N):
    try:
        t = open("$out/bench_n%d.json" % n).read()
        d = json.loads(t[t.index('{"metric'):])
        print("N=%d value %.1f ms %.4f e2e %s max_rel_err %s" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("grad_allreduce") or {}).get("max_rel_err")))
        for k in ("c4", "c5"):
            if k in d: print("  ", k, json.dumps({x: d[k].get(x) for x in ("value", "unit", "ms_per_step", "sweep", "error", "steps_redone_after_overflow")}))
        if d.get("extras_error"): print("  extras_error", d["extras_error"])
    except Exception as e:
        print("N=%d failed" % n, e); print(open("$out/bench_n%d.err" % n).read()[-2500:])
PY
