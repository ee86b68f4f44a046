#!/bin/bash
# end-of-round pass on one B200: DRAM traffic of two consecutive stages (metrics-only ncu pass), then the full GPU test suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02_final3.log
: > $L
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:euler -s 20 -c 4 --csv --log-file gpurun_out/traffic_r02g.csv \
   python bench.py --steps 2 --warmup 1 --min-time 0.01 --no-cpu --no-advection --e2e-steps 0 > /dev/null 2>&1
echo "traffic rc $?" >> $L
( time timeout 1500 python -m pytest tests -q -m gpu -x --durations=5 ) >> $L 2>&1
echo "pytest rc $?" >> $L
tail -25 $L
