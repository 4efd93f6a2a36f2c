"""A/B of hr_render's sub-batch size (hr_set_sub_batch) on the bench workload: ms per step, CUDA events, L2 flushed."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

hb, cfg, ds, sig, sd = bench.build_workload()
dev = torch.device("cuda", 0)
flush = torch.empty(bench.L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
res = []
for n in (65536, 1048576):
    rays = hb.rays.for_signature(sig, n, seed=5).to(dev)
    ref = None
    for sub in (-1, 0, 2 * 148 * 128, 4 * 148 * 128):
        model, render = bench.make_render(hb, cfg, ds, sd)
        model.set_sub_batch(sub)
        step = lambda: render(rays)["rgb"]  # noqa: E731
        for _ in range(5):
            out = step()
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        assert torch.equal(out, ref)
        ms = bench.timed_steps(torch, step, 20, flush) / 20
        tm = bench.kernel_times(torch, model, step, 10, flush)
        res.append({"rays": n, "sub_batch": sub, "ms_per_step": ms, "Mrays_s": n / ms / 1e3, "mlp_ms": tm["mlp_ms"], "render_ms": tm["render_ms"],
                    "workspace_MB": int(model._lib.hr_workspace_bytes(model._handle, n)) / 1e6})
        print(res[-1], flush=True)
        del model, render
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ab_subbatch.json"), "w"), indent=1)
