#!/bin/bash
# quick GPU check: parity subset + short bench with per-warp profile
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
MGK_PROFILE=1 timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/$1.log 2>&1
grep -v "^{" gpurun_out/$1.log | awk "NR%3==1"
