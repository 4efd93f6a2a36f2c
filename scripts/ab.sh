# Same-box A/B of two builds of the library (ab/A.so, ab/B.so; git-ignored, they travel with the gpurun snapshot):
# box-to-box spread of the same binary is +-3 %, more than most kernel changes, so variants are compared inside one call.
mkdir -p gpurun_out
for v in $(ls ab | sed "s/.so//") $(ls ab | sed "s/.so//"); do
  cp ab/$v.so hyperreel_b200/libhyperreel_b200.so
  if [ ! -f gpurun_out/ab_$v.parity ]; then  # a fast wrong kernel is not a result: rgb goldens + every shipped YAML first
    timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_shipped_yaml_gpu.py tests/test_parity_bites_gpu.py -m gpu -q -x -k "rgb_matches or fp32_path or per_sample or appearance_errors or host" > gpurun_out/ab_$v.parity 2>&1
    echo "$v parity: $(tail -1 gpurun_out/ab_$v.parity)"
  fi
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
d = json.loads(open(f'gpurun_out/ab_{v}.json').read().strip().splitlines()[-1])
x = d['extra_workloads']
print(v, 'T value %.1f e2e %.1f render %.4f net %.4f | D render %.4f | N3 render %.4f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['sample_net_kernel_ms'],
      x['donerf_sphere_s16']['roofline']['kernel_ms'], x['neural3d_s64']['roofline']['kernel_ms']))
PY
done
