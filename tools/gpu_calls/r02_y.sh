#!/bin/bash
# round 2 call y (2 GPUs): deferred peer wait on the owner of the next block (multi-GPU chain): dist tests, c4 bench with / without
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_dist.py -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee $O/r02y_pytest_dist.log
for o in "" "defer_wait=0"; do
  tag=$(echo "$o" | tr -d ' ' | tr '=' '_'); [ -z "$tag" ] && tag=default
  if [ -z "$o" ]; then optarg=""; else optarg="--opt $o"; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 $optarg 2>&1 | tail -1 > $O/r02y_bench_c4_2gpu_$tag.log
  python - "$O/r02y_bench_c4_2gpu_$tag.log" "c4 2gpu $tag" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step %.2f" % d["ms_per_step"], {k: round(v,2) for k,v in d["phases_ms"].items()}, d["multi_gpu_check"]["max_rel_dev_mean_vs_1gpu"])
except Exception as e:
    print(sys.argv[2], "FAILED", open(sys.argv[1]).read()[-400:])
PY
done | tee $O/r02y_bench_summary.txt
for w in c2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload $w --steps 10 --warmup 3 2>&1 | tail -1 > $O/r02y_bench_${w}_2gpu.log
  python - "$O/r02y_bench_${w}_2gpu.log" "$w 2gpu" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step %.2f" % d["ms_per_step"], {k: round(v,2) for k,v in d["phases_ms"].items()}, d["multi_gpu_check"]["max_rel_dev_mean_vs_1gpu"])
except Exception as e:
    print(sys.argv[2], "FAILED", open(sys.argv[1]).read()[-400:])
PY
done | tee -a $O/r02y_bench_summary.txt
