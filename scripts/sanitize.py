"""Small renders of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
    compute-sanitizer --tool memcheck python scripts/sanitize.py
Covers: tensor-core sample net (one and two input chunks, width 128 and 256, ragged last tile, several tiles per CTA),
fp32 sample net, render kernel variants ([8,0,0] / [8,4,4], S <= 32 and S = 64, SH and RGB shading, EXTRA outputs),
ray generation, the uint8 epilogue, the host-buffer pipeline."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hyperreel_b200 as hb  # noqa: E402
from tests.cases import FIELD_KWARGS, build_case  # noqa: E402

N = int(os.environ.get("HR_SANITIZE_RAYS", "700"))
for name, mode in (("technicolor_app", "auto"), ("neural3d_app", "auto"), ("donerf_wide_pe", "auto"), ("shiny_tiny", "auto"),
                   ("donerf_app", "fp32"), ("immersive_sphere_new", "fp32")):
    case = build_case(name, n=N)
    model = hb.LightfieldModel(case.model_cfg, dataset=case.dataset, mlp_mode=mode)
    render = hb.RenderLightfield(model, None, case.model_cfg.render)
    render.load_state_dict(case.state_dict, strict=False)
    render.eval()
    rays = case.rays.cuda()
    rgb = render(rays)["rgb"]
    st = model.render_stages(rays)
    out = render(rays, **FIELD_KWARGS)
    emb = render.embed(rays)
    u8 = model.render_to8b(rays)
    host = model.render_host(case.rays.clone().pin_memory())
    torch.cuda.synchronize()
    assert torch.equal(host, rgb.cpu()) and torch.isfinite(rgb).all()
    print(name, mode, "ok", float(rgb.mean()))
cam = hb.Camera(pose=[[1, 0, 0, 0.0], [0, 1, 0, 0.0], [0, 0, 1, 0.0]], K=[[60, 0, 32], [0, 60, 24], [0, 0, 1]], width=64, height=48, time=0.5)
print("rays", tuple(hb.generate_rays(cam, 8).shape))
torch.cuda.synchronize()
print("sanitize run complete")
