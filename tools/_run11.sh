python -m pytest tests/test_gpu_async_transfers.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-advection > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err; tail -2 gpurun_out/r02_bench5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench5.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e'])
PY
