#!/bin/bash
# last GPU call of round 1 (1 GPU, ~20 s): the notebook-parity test on the device + the find_MAP device/host comparison.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_notebook_parity.py tests/test_gpu_parity.py -m gpu -q -k "notebook or find_map" 2>&1 | tail -5 | tee gpurun_out/pytest_notebook_r01.log
