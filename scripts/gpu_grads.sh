mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_grads_gpu.py -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_grads.log 2>&1; echo "grads rc=$?"; grep -E "^(FAILED|ERROR)|Error|assert |passed|failed" gpurun_out/pytest_grads.log | head -40
timeout 900 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_1gpu.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print(json.dumps(d['extra_workloads'].get('train_step'), indent=1))
PY
