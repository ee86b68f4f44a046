#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02_var3.log
: > $L
run() {
  python bench.py --order $1 --no-cpu --no-advection --min-time 1 --steps 20 --e2e-steps 2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()[-300:]); continue
    print(d['value'], d['roofline']['kernel_ms'], d['fp64']['frac'])
" >> $L 2>&1
}
for V in base t448 t384 t384e f448; do
  for N in 4 3; do
    echo "== $V N=$N" >> $L
    if [ $V = base ]; then run $N; else HDG_LIB_PATH=hopefoam_b200/variants/libhopedg_$V.so run $N; fi
  done
done
tail -40 $L
