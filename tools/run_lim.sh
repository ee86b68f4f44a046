#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02_lim4.log
: > $L
for C in 0 8; do
HDG_LIM_CFG=$C timeout 600 python -m pytest tests/test_gpu_limiter.py -x -q -m gpu >> $L 2>&1
echo "pytest cfg $C rc $?" >> $L
done
for C in 0 8; do
  echo "== HDG_LIM_CFG=$C" >> $L
  HDG_LIM_CFG=$C python tests/perf_limiter.py 500 4 >> $L 2>&1
  HDG_LIM_CFG=$C python tests/perf_limiter.py 707 4 >> $L 2>&1
  HDG_LIM_CFG=$C timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lim -s 9 -c 3 python tests/perf_limiter.py 500 4 2>&1 | grep -E "void|duration" >> $L
done
tail -50 $L
