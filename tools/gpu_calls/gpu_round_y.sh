#!/bin/bash
# GPU call y (2 GPUs): the four multi-process modes after the stream-choreography change + one sharded bench point.
TAG=${1:-r01y}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_dist.py -m gpu -q 2>&1 | tail -4 | tee $O/pytest_dist_$TAG.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_c2_fp64_g2_$TAG.json 2> $O/bench_c2_fp64_g2_$TAG.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_c2_fp64_g2_$TAG.json") if l.startswith("{")][-1])
    print("c2 2gpu", {k: d[k] for k in ("value", "ms_per_step", "phases_ms")}, "e2e", d["e2e"]["value"])
except Exception as e:
    print("bench failed", e); print(open("$O/bench_c2_fp64_g2_$TAG.err").read()[-2500:])
PY
