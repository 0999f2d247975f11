#!/bin/bash
TAG=${1:-r01m}
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest_gpu_$TAG.log
echo "== find_MAP timing"; timeout 600 python tools/map_timing.py 2048 2>&1 | tail -3 | tee $O/map_timing_$TAG.log; timeout 600 python tools/map_timing.py 8192 2>&1 | tail -3 | tee -a $O/map_timing_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
