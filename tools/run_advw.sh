#!/bin/bash
# wide-row TMA advection kernel: parity tests, then the configuration sweep (N = 5, 6)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02_advw.log
: > $L
timeout 600 python -m pytest tests/test_gpu_advection.py -x -q -m gpu -k "not large_mesh and not alternate and not config1 and not 1000 and (5 or 6)" >> $L 2>&1
echo "pytest rc $?" >> $L
for N in 5 6; do
  for C in 0 1 2 3; do
    echo "== N=$N cfg=$C" >> $L
    HDG_ADVW_CFG=$C timeout 120 python bench.py --workload advection --order $N --steps 100 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()[-300:]); continue
    print(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['kernel'])
" >> $L 2>&1
  done
done
tail -40 $L
