#!/bin/bash
# round 2 call o (2 GPUs): multi-GPU tests (sharded factorisation, lock-step find_MAP, device all-gather in predict) + c4 bench on 2 GPUs, both arms
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | tee $O/r02o_gpus.log
timeout 900 python -m pytest tests/test_dist.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 | tee $O/r02o_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 | tee $O/r02o_bench_c4_2gpu.log | cut -c1-2500
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 ) 2>&1 | tail -5 | tee $O/r02o_bench_reference_2gpu.log | cut -c1-1500
