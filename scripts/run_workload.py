"""Render one of bench.py's workloads a few times (for ncu captures of the off-headline configurations).
    python scripts/run_workload.py donerf_sphere_s16|neural3d_s64|headline [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

name = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
if name == "headline":
    hb, cfg, ds, sig, sd = bench.build_workload()
    n = bench.RAYS_PER_GPU
else:
    spec = bench.EXTRA_WORKLOADS[name]
    hb, cfg, ds, sig, sd = bench.build_workload(spec["builtin"], spec["over"], gain=100.0, app_gain=6.0)
    n = spec["rays"]
model, render = bench.make_render(hb, cfg, ds, sd)
rays = hb.rays.for_signature(sig, n, seed=5).cuda()
flush = torch.empty(bench.L2_FLUSH_BYTES // 4, dtype=torch.float32, device="cuda")
for _ in range(iters):
    flush.zero_()
    out = render(rays)["rgb"]
torch.cuda.synchronize()
print(name, tuple(out.shape), float(out.mean()))
