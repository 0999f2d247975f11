#!/bin/bash
# GPU call k (N GPUs): sharded factorisation check + strong-scaling benches.
TAG=${1:-r01k}
NG=${2:-4}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
echo "== dist_check p2p fp64";  timeout 300 $TR --master-port 29511 tests/dist_check.py 1000 3000 8192 16384 2>&1 | grep -E "^\{|DIST_CHECK|rror|Traceback" | tee $O/dist_check_p2p_g${NG}_$TAG.log
echo "== dist_check p2p tf32";  GB2_DIST_PRECISION=tf32 timeout 300 $TR --master-port 29513 tests/dist_check.py 3000 8192 2>&1 | grep -E "^\{|DIST_CHECK|rror|Traceback" | tee $O/dist_check_p2p_tf32_g${NG}_$TAG.log
for cfg in "c2 fp64" "c4 fp64" "c4 tf32"; do set -- $cfg
  echo "== bench --gpus $NG $1 $2"; timeout 900 $TR --master-port 29514 bench.py --gpus $NG --workload $1 --precision $2 --steps 3 --warmup 3 > $O/bench_$1_$2_g${NG}_$TAG.json 2> $O/bench_$1_$2_g${NG}_$TAG.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_$1_$2_g${NG}_$TAG.json") if l.startswith("{")][-1])
    print("$1 $2", {k: d[k] for k in ("value", "ms_per_step", "phases_ms", "cholesky_tflops")}, "e2e", d["e2e"]["value"], "warm", d["warm"]["ms_per_step"])
except Exception as e:
    print("bench failed", e); print(open("$O/bench_$1_$2_g${NG}_$TAG.err").read()[-2500:])
PY
done
