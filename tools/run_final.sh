#!/bin/bash
# end-of-round measurement pass on one B200: full GPU test suite, advection order sweep, ncu captures, bench line + launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02_final.log
: > $L
( time timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 ) >> $L 2>&1
echo "pytest rc $?" >> $L
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
for N in 4 5 6; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:advectStageTma -s 6 -c 1 -f -o gpurun_out/prof_adv_r02f_N$N \
     python bench.py --workload advection --order $N --steps 5 > /dev/null 2>&1
  echo "ncu adv N=$N rc $?" >> $L
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lim -s 9 -c 3 -f -o gpurun_out/prof_limiter_r02f python tests/perf_limiter.py 500 4 > /dev/null 2>&1
echo "ncu limiter rc $?" >> $L
python tests/perf_limiter.py 500 4 >> $L 2>&1
python tests/perf_limiter.py 707 4 >> $L 2>&1
python tests/perf_limiter.py 500 6 >> $L 2>&1
timeout 600 python bench.py > gpurun_out/bench_1gpu_r02f.json 2>> $L
echo "bench rc $?" >> $L
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r02f.json 2>> $L
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02f.csv python bench.py --steps 2 --warmup 1 --min-time 0.01 --no-cpu > /dev/null 2>&1
echo "launch list rc $?" >> $L
tail -60 $L
