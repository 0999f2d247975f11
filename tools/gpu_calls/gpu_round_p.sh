#!/bin/bash
# GPU call p (8 GPUs): storage-sharded mode at 8 ranks, then BASELINE config 5 (stacked N = 262144).
TAG=${1:-r01p}
NG=${2:-8}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
echo "== dist_check shard_storage"; GB2_DIST_SHARD=1 timeout 300 $TR --master-port 29511 tests/dist_check.py 3000 8192 2>&1 | grep -E "^\{|DIST_CHECK|rror|Traceback|rap|ssert" | tee $O/dist_check_shard_g${NG}_$TAG.log
if ! grep -q "DIST_CHECK OK" $O/dist_check_shard_g${NG}_$TAG.log; then echo "dist_check failed: skipping the large run"; exit 1; fi
echo "== c5 demo small (n=8192 -> N=32768)"; timeout 300 $TR --master-port 29512 tools/c5_demo.py 8192 1 2>&1 | grep -E "^\{|RESULT|rror|Traceback" | cut -c1-1500 | tee $O/c5_demo_small_$TAG.log
if ! grep -q "RESULT_IDENTICAL_ON_ALL_RANKS True" $O/c5_demo_small_$TAG.log; then echo "small demo failed: skipping the large run"; exit 1; fi
echo "== c5 demo full (n=65536 -> N=262144)"; timeout 600 $TR --master-port 29513 tools/c5_demo.py 65536 1 2>&1 | grep -E "^\{|RESULT|rror|Traceback" | cut -c1-1500 | tee $O/c5_demo_full_$TAG.log
nvidia-smi --query-gpu=index,memory.used --format=csv | head -9
