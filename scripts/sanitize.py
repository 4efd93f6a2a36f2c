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
# round 2 additions: 96 / 128 / 256 samples per ray (4 / 8 samples per lane), voxel grids, the colour transform, the cascaded
# pipelines (cascade_points_kernel, point net on point rows), through the shipped-YAML fixtures; the backward kernel
from tests.test_shipped_yaml_golden import SHIPPED, load_fixture  # noqa: E402

BY_NAME = {os.path.basename(p)[:-4]: p for p in SHIPPED}
for name, mode in (("neural_3d_z_plane_static", "auto"), ("technicolor_z_plane_no_sample", "auto"), ("catacaustics_voxel", "auto"),
                   ("shiny_z_deformable", "fp32"), ("immersive_z_plane", "auto"), ("technicolor_cascaded", "auto"),
                   ("shiny_z_plane_cascaded", "fp32"), ("shiny_z_tensorf_cascaded", "auto")):
    plain, cfg, ds, sig, sd, rays, rgb_ref = load_fixture(BY_NAME[name])
    model = hb.LightfieldModel(cfg, dataset=ds, mlp_mode=mode)
    render = hb.RenderLightfield(model, None, cfg.render)
    render.load_state_dict(sd, strict=False)
    render.eval()
    dev = rays.repeat(4, 1)[:333].contiguous().cuda()
    rgb = render(dev)["rgb"]
    st = model.render_stages(dev)
    emb = render.embed(dev)
    host = model.render_host(dev.cpu().pin_memory())
    torch.cuda.synchronize()
    assert torch.equal(host, rgb.cpu()) and float((rgb[:rays.shape[0]].cpu() - rgb_ref[:333]).abs().max()) <= 1e-4
    print(name, mode, "ok", float(rgb.mean()))
for name in ("technicolor_app", "donerf_app"):
    case = build_case(name, n=300)
    model = hb.LightfieldModel(case.model_cfg, dataset=case.dataset)
    render = hb.RenderLightfield(model, None, case.model_cfg.render)
    render.load_state_dict(case.state_dict, strict=False)
    render.cuda().train()
    out = model.render_differentiable(case.rays.cuda(), white_bg=False)
    out.square().sum().backward()
    torch.cuda.synchronize()
    print(name, "backward ok", float(out.mean()))
cam = hb.Camera(pose=[[1, 0, 0, 0.0], [0, 1, 0, 0.0], [0, 0, 1, 0.0]], K=[[60, 0, 32], [0, 60, 24], [0, 0, 1]], width=64, height=48, time=0.5)
print("rays", tuple(hb.generate_rays(cam, 8).shape))
torch.cuda.synchronize()
print("sanitize run complete")
