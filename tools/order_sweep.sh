#!/bin/bash
# Order sweep of the Euler stage on one GPU (run on the GPU box):  tools/order_sweep.sh "1 2 3 4" [label] [extra bench args]
# prints one line per order with the stage time and the FP64 / HBM roofline fractions; env (HDG_EULER_SPLIT, HDG_LIB_PATH) passes through
orders=${1:-"1 2 3 4 5 6 7 8"}; label=${2:-sweep}; shift 2 || true
for N in $orders; do
  python bench.py --order $N --no-cpu --no-advection --steps 10 --warmup 3 --min-time 0.6 --e2e-steps 4 --e2e-serial "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']; f=d['fp64']
print('$label euler N=$N %.2f GDOF/s stage_ms %.4f hbm_frac %.3f fp64_frac %.3f launches/step %d' % (d['value'], r['kernel_ms'], r['frac'], f['frac'], d['gpu_launches']/d['steps']))
"
done
