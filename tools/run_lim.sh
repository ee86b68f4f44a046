#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02_lim5.log
: > $L
HDG_LIM_CFG=16 timeout 600 python -m pytest tests/test_gpu_limiter.py -x -q -m gpu 2>&1 | tail -1 >> $L
for C in 0 16; do
  echo "== HDG_LIM_CFG=$C" >> $L
  HDG_LIM_CFG=$C python tests/perf_limiter.py 500 4 >> $L 2>&1
  HDG_LIM_CFG=$C python tests/perf_limiter.py 707 4 >> $L 2>&1
done
tail -50 $L
