#!/bin/bash
# compute-sanitizer pass over the two stage kernels (run under gpurun): memcheck + racecheck + initcheck on small cases.
set -x
for tool in memcheck racecheck initcheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_euler_stage.py -q -m gpu -k "ragged or periodic or wall" -x 2>&1 | tail -4
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_advection.py -q -m gpu -k "periodic_and_zero" -x 2>&1 | tail -4
done
