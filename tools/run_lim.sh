#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02_lim2.log
: > $L
timeout 600 python -m pytest tests/test_gpu_limiter.py -x -q -m gpu >> $L 2>&1
echo "pytest rc $?" >> $L
for S in 0 1; do
  echo "== HDG_LIM_STREAM=$S" >> $L
  HDG_LIM_STREAM=$S python tests/perf_limiter.py 500 4 >> $L 2>&1
  HDG_LIM_STREAM=$S python tests/perf_limiter.py 707 4 >> $L 2>&1
  HDG_LIM_STREAM=$S timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:lim -s 9 -c 3 python tests/perf_limiter.py 500 4 2>&1 | grep -E "void|duration|dram" >> $L
done
tail -50 $L
