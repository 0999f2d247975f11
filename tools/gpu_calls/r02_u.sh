#!/bin/bash
# round 2 call u (8 GPUs): c4 fp64 bench on 8 and on 4 GPUs (sharded factorisation with two-level blocking + TMA-staged GEMM, exactness check inside)
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | wc -l | tee $O/r02u_gpus.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 2>&1 | tail -1 | tee $O/r02u_bench_c4_8gpu.log | cut -c1-1200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 5 --warmup 3 2>&1 | tail -1 | tee $O/r02u_bench_c4_4gpu.log | cut -c1-1200
