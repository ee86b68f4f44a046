#!/bin/bash
# end-of-round measurement pass on one B200: bench line + reference arm + launch list, Euler order sweep, ncu captures of the stage kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02_final2.log
: > $L
timeout 600 python bench.py > gpurun_out/bench_1gpu_r02g.json 2>> $L
echo "bench rc $?" >> $L
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r02g.json 2>> $L
bash tools/order_sweep.sh "1 2 3 4 5 6 7 8 9 10" final >> $L 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02g.csv python bench.py --steps 2 --warmup 1 --min-time 0.01 --no-cpu > /dev/null 2>&1
echo "launch list rc $?" >> $L
timeout 300 ncu --set full --clock-control none --import-source on -k regex:euler -s 20 -c 2 -f -o gpurun_out/prof_euler_r02g python bench.py --steps 2 --warmup 1 --min-time 0.01 --no-cpu --no-advection --e2e-steps 0 > /dev/null 2>&1
echo "ncu euler rc $?" >> $L
tail -30 $L
