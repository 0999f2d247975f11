#!/bin/bash
# GPU call v (2 GPUs): final multi-GPU regression incl. the MLL gradient on a sharded factorisation + c3 scaling point.
TAG=${1:-r01v}
NG=${2:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
echo "== dist_check p2p fp64 (+ gradient)"; timeout 300 $TR --master-port 29511 tests/dist_check.py 1000 3000 8192 2>&1 | grep -E "^\{|DIST_CHECK|rror|Traceback" | tee $O/dist_check_p2p_g${NG}_$TAG.log
echo "== pytest tests/test_dist.py -m gpu"; timeout 900 python -m pytest tests/test_dist.py -m gpu -q 2>&1 | tail -4
echo "== bench c2 --gpus $NG"; timeout 600 $TR --master-port 29514 bench.py --gpus $NG --steps 5 --warmup 3 > $O/bench_c2_fp64_g${NG}_$TAG.json 2> $O/bench_c2_fp64_g${NG}_$TAG.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_c2_fp64_g${NG}_$TAG.json") if l.startswith("{")][-1])
    print("c2", {k: d[k] for k in ("value", "ms_per_step", "phases_ms")}, "e2e", d["e2e"]["value"])
except Exception as e:
    print("bench failed", e); print(open("$O/bench_c2_fp64_g${NG}_$TAG.err").read()[-2500:])
PY
