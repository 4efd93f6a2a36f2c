mkdir -p gpurun_out
bash scripts/gpu_check.sh
timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/r2m_bench_1gpu.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2m_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['sample_net_kernel_ms'])
for k,v in d['extra_workloads'].items():
    if 'roofline' in v: print(k, v['value'], v['ms_per_step'], v['roofline']['kernel_ms'], v['roofline']['sample_net_kernel_ms'], v['roofline']['frac'])
PY
