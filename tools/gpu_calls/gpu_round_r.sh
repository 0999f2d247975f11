#!/bin/bash
# GPU call r (2 GPUs): the whole -m gpu suite including the multi-process tests (p2p / nccl / shard_storage / tf32).
TAG=${1:-r01r}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $O/pytest_gpu_2gpu_$TAG.log
