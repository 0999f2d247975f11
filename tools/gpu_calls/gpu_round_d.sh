#!/bin/bash
# GPU call d (2+ GPUs): NCCL row-block-sharded Cholesky vs single GPU, then the sharded bench.
TAG=${1:-r01d}
NG=${2:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi -L | tee $O/gpus_$TAG.txt
echo "== dist_check"; NCCL_DEBUG=WARN timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py 1000 3000 8192 16384 2>&1 | grep -v "^W" | tail -30 | tee $O/dist_check_$TAG.log
echo "== bench --gpus $NG (c2 fp64)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 5 --warmup 3 > $O/bench_c2_g${NG}_$TAG.json 2> $O/bench_c2_g${NG}_$TAG.err; tail -c 1500 $O/bench_c2_g${NG}_$TAG.json; tail -5 $O/bench_c2_g${NG}_$TAG.err
echo "== bench --gpus $NG (c4 fp64)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --workload c4 --steps 3 --warmup 3 > $O/bench_c4_g${NG}_$TAG.json 2> $O/bench_c4_g${NG}_$TAG.err; tail -c 1500 $O/bench_c4_g${NG}_$TAG.json; tail -5 $O/bench_c4_g${NG}_$TAG.err
