#!/bin/bash
# The measurement pass behind profiles/*_r02b.* (run on a B200 box:  gpurun --timeout 2400 -- 'bash tools/run_final.sh').
# Everything lands in gpurun_out/; the summaries under profiles/ are made from it here (tools/ncu_summary.py, tools/ncu_brief.py).
#   1. bench line + reference arm                      -> bench_1gpu.json, bench_ref.json
#   2. order sweeps: Euler N = 1..10, advection N = 1..8 -> final.log
#   3. ncu launch list of the bench command            -> launches.csv            (shares of the step)
#   4. ncu --set full: one Euler stage, advection N = 4, 5, 6, the limiter's three kernels -> prof_*.ncu-rep
#   5. DRAM traffic of two consecutive Euler stages    -> traffic.csv             (profiles/ncu_traffic_r02b.json, read by bench.py)
#   6. limiter timing, full GPU test suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/final.log
: > $L
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2>> $L
echo "bench rc $?" >> $L
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> $L
bash tools/order_sweep.sh "1 2 3 4 5 6 7 8 9 10" final >> $L 2>&1
for N in 1 2 3 4 5 6 7 8; do
  timeout 120 python bench.py --workload advection --order $N --steps 100 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']
    print('final advection N=$N %.2f GDOF/s kernel_ms %.4f hbm_frac %.3f %s' % (d['value'], r['kernel_ms'], r['frac'], r['kernel']))
" >> $L 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --min-time 0.01 --no-cpu > /dev/null 2>&1
echo "launch list rc $?" >> $L
B="python bench.py --steps 2 --warmup 1 --min-time 0.01 --no-cpu --no-advection --e2e-steps 0"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:euler -s 20 -c 2 -f -o gpurun_out/prof_euler $B > /dev/null 2>&1
echo "ncu euler rc $?" >> $L
for N in 4 5 6; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:advectStageTma -s 6 -c 1 -f -o gpurun_out/prof_adv_N$N \
     python bench.py --workload advection --order $N --steps 5 > /dev/null 2>&1
  echo "ncu advection N=$N rc $?" >> $L
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lim -s 9 -c 3 -f -o gpurun_out/prof_limiter python tests/perf_limiter.py 500 4 > /dev/null 2>&1
echo "ncu limiter rc $?" >> $L
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:euler -s 20 -c 4 --csv --log-file gpurun_out/traffic.csv $B > /dev/null 2>&1
echo "traffic rc $?" >> $L
python tests/perf_limiter.py 500 4 >> $L 2>&1
python tests/perf_limiter.py 707 4 >> $L 2>&1
python tests/perf_limiter.py 500 6 >> $L 2>&1
( time timeout 1500 python -m pytest tests -q -m gpu -x --durations=5 ) >> $L 2>&1
echo "pytest rc $?" >> $L
tail -60 $L
