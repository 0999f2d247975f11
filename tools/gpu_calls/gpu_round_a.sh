#!/bin/bash
# One gpurun call: parity tests, bench (ours + reference arm), ncu launch list, ncu --set full of the top kernels, library peaks.
# Usage (from the repo root, on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu_$TAG.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_gpu_$TAG.log
echo "== bench ours" ; timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_c2_$TAG.json 2> $O/bench_c2_$TAG.err; tail -c 3000 $O/bench_c2_$TAG.json; tail -5 $O/bench_c2_$TAG.err
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_c2_$TAG.json 2> $O/bench_ref_$TAG.err; cat $O/bench_ref_c2_$TAG.json
echo "== peaks" ; timeout 600 python tools/peaks_torch.py 2>&1 | tail -2
echo "== ncu launch list" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/bench_under_ncu_$TAG.log 2>&1
python tools/launch_summary.py $O/launches_$TAG.csv | tee $O/launch_summary_$TAG.txt | head -20
echo "== ncu full" 
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_nt -s 300 -c 4 -o $O/prof_dgemm_$TAG -f python tools/prof_factorize.py 8192 1 > $O/ncu_dgemm_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kbuild -c 2 -o $O/prof_kbuild_$TAG -f python tools/prof_factorize.py 8192 1 > $O/ncu_kbuild_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_diag -s 3 -c 1 -o $O/prof_potrf_$TAG -f python tools/prof_factorize.py 8192 1 > $O/ncu_potrf_$TAG.log 2>&1
ls -la $O
