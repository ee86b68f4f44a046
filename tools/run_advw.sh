#!/bin/bash
# TMA advection kernels: parity tests, then the configuration sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02_advw3.log
: > $L
timeout 600 python -m pytest tests/test_gpu_advection.py -x -q -m gpu -k "not large_mesh and not alternate and not wide_kernel and not config1 and not 1000" >> $L 2>&1
echo "pytest rc $?" >> $L
run() {
    python bench.py --workload advection --order $1 --steps 100 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()[-300:]); continue
    print(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['kernel'])
" >> $L 2>&1
}
for C in 1 3 4 5; do echo "== N=4 HDG_ADV_CFG=$C" >> $L; HDG_ADV_CFG=$C run 4; done
for C in 1 3; do echo "== N=3 HDG_ADV_CFG=$C" >> $L; HDG_ADV_CFG=$C run 3; done
for N in 1 2; do
  for C in 0 1 2 3; do echo "== N=$N cfg=$C" >> $L; HDG_ADVW_CFG=$C run $N; done
done
tail -40 $L
