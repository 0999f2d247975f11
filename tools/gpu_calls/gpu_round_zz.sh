#!/bin/bash
# GPU call zz (2 GPUs): collective agreement on the factorisation verdict (non-PD case) + regression.
TAG=${1:-r01zz}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29511 tests/dist_check.py 1000 3000 2>&1 | grep -E "^\{|DIST_CHECK|rror|Traceback" | cut -c1-400 | tee $O/dist_check_p2p_g2_$TAG.log
GB2_DIST_SHARD=1 timeout 120 $TR --master-port 29512 tests/dist_check.py 1000 2>&1 | grep -E "non_pd|DIST_CHECK|rror|Traceback" | cut -c1-400 | tee $O/dist_check_shard_g2_$TAG.log
