# Same-box A/B of two builds of the library (ab/A.so, ab/B.so; git-ignored, they travel with the gpurun snapshot):
# box-to-box spread of the same binary is +-3 %, more than most kernel changes, so variants are compared inside one call.
mkdir -p gpurun_out
for v in $(ls ab | sed "s/.so//") $(ls ab | sed "s/.so//"); do
  cp ab/$v.so hyperreel_b200/libhyperreel_b200.so
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
d = json.loads(open(f'gpurun_out/ab_{v}.json').read().strip().splitlines()[-1])
x = d['extra_workloads']
print(v, 'T value %.1f render %.4f net %.4f | D render %.4f | N3 render %.4f' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['sample_net_kernel_ms'],
      x['donerf_sphere_s16']['roofline']['kernel_ms'], x['neural3d_s64']['roofline']['kernel_ms']))
PY
done
