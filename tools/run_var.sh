#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02_var2.log
: > $L
run() {
  python bench.py --order $1 --no-cpu --no-advection --min-time 1 --steps $2 --e2e-steps 2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()[-300:]); continue
    print(d['value'], d['roofline']['kernel_ms'], d['fp64']['frac'])
" >> $L 2>&1
}
timeout 600 python -m pytest tests/test_gpu_euler_split.py tests/test_gpu_euler_stage.py tests/test_gpu_halo_parity.py tests/test_gpu_euler_run.py -x -q -m gpu 2>&1 | tail -2 >> $L
for N in 1 2 3 4 5 6 7 8; do echo "== base N=$N" >> $L; run $N 20; done
for N in 5 6; do echo "== old56 N=$N" >> $L; HDG_LIB_PATH=hopefoam_b200/variants/libhopedg_old56.so run $N 20; done
for N in 7 8; do echo "== oldface7 N=$N" >> $L; HDG_LIB_PATH=hopefoam_b200/variants/libhopedg_oldface7.so run $N 20; done
for N in 1 4; do echo "== oldlow N=$N" >> $L; HDG_LIB_PATH=hopefoam_b200/variants/libhopedg_oldlow.so run $N 20; done
tail -60 $L
