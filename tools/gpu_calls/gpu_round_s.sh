#!/bin/bash
# GPU call s (2 GPUs): tf32 panels on the storage-sharded layout vs single GPU; regression of the other sharded modes.
TAG=${1:-r01s}
NG=${2:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
echo "== dist_check shard_storage tf32"; GB2_DIST_SHARD=1 GB2_DIST_PRECISION=tf32 timeout 300 $TR --master-port 29511 tests/dist_check.py 3000 8192 16384 2>&1 | grep -E "^\{|DIST_CHECK|rror|Traceback|rap|ssert" | tee $O/dist_check_shard_tf32_g${NG}_$TAG.log
echo "== dist_check shard_storage fp64"; GB2_DIST_SHARD=1 timeout 300 $TR --master-port 29512 tests/dist_check.py 3000 8192 2>&1 | grep -E "^\{|DIST_CHECK|rror|Traceback" | tee $O/dist_check_shard_g${NG}_$TAG.log
echo "== test_random_models"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "random_models" 2>&1 | tail -3
echo "== bench c4 tf32 shard"; timeout 900 $TR --master-port 29514 bench.py --gpus $NG --workload c4 --precision tf32 --steps 3 --warmup 3 --opt shard_storage=1 > $O/bench_c4_tf32_shard_g${NG}_$TAG.json 2> $O/bench_c4_tf32_shard_g${NG}_$TAG.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_c4_tf32_shard_g${NG}_$TAG.json") if l.startswith("{")][-1])
    print("c4 tf32 shard", {k: d[k] for k in ("value", "ms_per_step", "phases_ms", "cholesky_tflops")}, "e2e", d["e2e"]["value"])
except Exception as e:
    print("bench failed", e); print(open("$O/bench_c4_tf32_shard_g${NG}_$TAG.err").read()[-2500:])
PY
