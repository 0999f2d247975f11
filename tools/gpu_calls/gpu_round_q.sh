#!/bin/bash
# GPU call q (1 GPU): diagonal-panel kernel phase clocks; full parity suite; default bench + reference arm; ncu launch list.
TAG=${1:-r01q}
O=gpurun_out
mkdir -p $O
echo "== micro_potrf"; timeout 60 tools/micro_potrf 2>&1 | tee $O/micro_potrf_$TAG.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_gpu_$TAG.log
echo "== bench default"; timeout 600 python bench.py > $O/bench_default_$TAG.json 2> $O/bench_default_$TAG.err; tail -c 1200 $O/bench_default_$TAG.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_under_ncu_$TAG.log 2>&1
python tools/launch_summary.py $O/launches_$TAG.csv | tee $O/launch_summary_$TAG.txt | head -14
