#!/bin/bash
# GPU call t (8 GPUs): BASELINE config 5 with tf32 (tcgen05) trailing updates on the storage-sharded layout.
TAG=${1:-r01t}
NG=${2:-8}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
echo "== c5 demo small tf32"; timeout 300 $TR --master-port 29512 tools/c5_demo.py 8192 1 tf32 2>&1 | grep -E "^\{|RESULT|rror|Traceback" | cut -c1-1500 | tee $O/c5_demo_small_tf32_$TAG.log
if ! grep -q "RESULT_IDENTICAL_ON_ALL_RANKS True" $O/c5_demo_small_tf32_$TAG.log; then echo "small demo failed: skipping the large run"; exit 1; fi
echo "== c5 demo full tf32"; timeout 600 $TR --master-port 29513 tools/c5_demo.py 65536 1 tf32 2>&1 | grep -E "^\{|RESULT|rror|Traceback" | cut -c1-1500 | tee $O/c5_demo_full_tf32_$TAG.log
