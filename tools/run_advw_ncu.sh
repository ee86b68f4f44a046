#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for N in 5 6; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:advectStageTmaWide -s 6 -c 1 -f -o gpurun_out/prof_advw_r02_N$N \
     python bench.py --workload advection --order $N --steps 5 > gpurun_out/prof_advw_N$N.log 2>&1
  echo "ncu N=$N rc $?"
done
