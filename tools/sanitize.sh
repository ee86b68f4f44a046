#!/bin/bash
# compute-sanitizer pass over the stage kernels (run under gpurun): memcheck + racecheck + initcheck + synccheck on small cases.
# The advection selection covers the TMA-pipelined kernels (128-B rows N = 3, 4; wide rows N = 1, 2, 5, 6, 7: periodic, zeroGradient, ragged
# octets, LSERK residual path, changing velocity); the limiter tests cover the three limiter kernels.
set -x
for tool in memcheck racecheck initcheck synccheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_euler_split.py tests/test_gpu_euler_stage.py -q -m gpu -k "split_equals_fused_all_orders or thin or ragged or periodic or wall or smallest" -x 2>&1 | tail -4
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_limiter.py -q -m gpu -x 2>&1 | tail -4
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_advection.py -q -m gpu -k "periodic_and_zero or ragged or smallest or lserk or velocity_changes or average" -x 2>&1 | tail -4
done
