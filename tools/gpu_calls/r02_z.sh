#!/bin/bash
# round 2 call z (8 GPUs): deferred peer wait at 8 GPUs (c4, with / without), C2 at 8 GPUs, and the lock-step distributed find_MAP check
mkdir -p gpurun_out
O=gpurun_out
for o in "" "defer_wait=0"; do
  tag=$(echo "$o" | tr -d ' ' | tr '=' '_'); [ -z "$tag" ] && tag=default
  if [ -z "$o" ]; then optarg=""; else optarg="--opt $o"; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 6 --warmup 3 $optarg 2>&1 | tail -1 > $O/r02z_bench_c4_8gpu_$tag.log
  python - "$O/r02z_bench_c4_8gpu_$tag.log" "c4 8gpu $tag" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step %.2f" % d["ms_per_step"], {k: round(v,2) for k,v in d["phases_ms"].items()}, "e2e %.2f" % d["e2e"]["ms_per_step"], d["multi_gpu_check"]["max_rel_dev_mean_vs_1gpu"])
except Exception as e:
    print(sys.argv[2], "FAILED", open(sys.argv[1]).read()[-400:])
PY
done | tee $O/r02z_bench_summary.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29553 tests/dist_check.py 1000 2>&1 | grep -E "find_MAP|DIST_CHECK|Error|error" | tail -5 | tee $O/r02z_dist_check_find_map.log
