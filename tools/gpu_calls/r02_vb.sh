#!/bin/bash
# round 2 call vb (2 GPUs): the sharded paths of the K-build v6 (row-block ownership, storage-sharded row mapping) -- dist tests, c4 bench on 2 GPUs
# with its in-bench comparison against the single-GPU result
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_dist.py -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee $O/r02vb_pytest_dist.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 > $O/r02vb_bench_c4_2gpu.log
python - "$O/r02vb_bench_c4_2gpu.log" "c4 2gpu" <<'PY' | tee $O/r02vb_bench_summary.txt
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step %.2f" % d["ms_per_step"], {k: round(v,2) for k,v in d["phases_ms"].items()}, d["multi_gpu_check"])
except Exception as e:
    print(sys.argv[2], "FAILED", open(sys.argv[1]).read()[-400:])
PY
