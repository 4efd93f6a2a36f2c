#!/usr/bin/env python
"""Where the end-to-end time of hr_render_host goes: copy times alone, and the host-buffer call at several chunk sizes."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hyperreel_b200 as hb
from hyperreel_b200.state import seeded_state_dict

dev = torch.device("cuda", 0)
cfg, ds = hb.configs.get("technicolor_z_plane", n_voxels=512000000)
sig = hb.lower(cfg, ds)
model = hb.LightfieldModel(cfg, dataset=ds, mlp_mode="bf16x3")
render = hb.RenderLightfield(model, None, cfg.render)
render.load_state_dict(seeded_state_dict(sig, seed=11, density_gain=30.0), strict=False)
render.eval()
n = 65536
rays_host = hb.rays.for_signature(sig, n, seed=5).pin_memory()
rgb_host = torch.empty((n, 3)).pin_memory()
rays_dev = rays_host.to(dev)
rgb_dev = torch.empty((n, 3), device=dev)
flush = torch.empty((512 << 20) // 4, dtype=torch.float32, device=dev)

def wall(fn, reps=30, do_flush=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return 1e3 * ts[len(ts) // 2]

out = {}
out["h2d_2MB_ms"] = wall(lambda: rays_dev.copy_(rays_host, non_blocking=True))
out["d2h_786KB_ms"] = wall(lambda: rgb_host.copy_(rgb_dev, non_blocking=True))
out["device_render_ms"] = wall(lambda: render(rays_dev))
out["render_host_default_ms"] = wall(lambda: model.render_host(rays_host, rgb_host))
out["render_host_default_noflush_ms"] = wall(lambda: model.render_host(rays_host, rgb_host), do_flush=False)
for chunk in (65536, 18944):
    out[f"render_host_chunk{chunk}_ms"] = wall(lambda: model.render_host(rays_host, rgb_host, chunk=chunk))
    out[f"render_host_chunk{chunk}_noflush_ms"] = wall(lambda: model.render_host(rays_host, rgb_host, chunk=chunk), do_flush=False)
print(json.dumps(out, indent=1))
