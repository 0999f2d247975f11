#!/bin/bash
# GPU call b: tcgen05 split-TF32 GEMM stand-alone check, new parity tests (MLL gradient, find_MAP), c4 (N=32768) bench.
TAG=${1:-r01b}
O=gpurun_out
mkdir -p $O
echo "== test_tf32"; timeout 180 tools/test_tf32 2>&1 | tee $O/test_tf32_$TAG.log
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_gpu_$TAG.log
echo "== bench c4 fp64"; timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 > $O/bench_c4_$TAG.json 2> $O/bench_c4_$TAG.err; tail -c 2500 $O/bench_c4_$TAG.json; tail -5 $O/bench_c4_$TAG.err
