"""Backward pass of the path on the GPU: gradients of every parameter against the reference's own autograd
(tests/golden/grads_*.npz, made by tests/golden/make_golden_grads.py from the unmodified reference) and against the gradient
oracle (training-mode forward: no clamp, white background)."""
import os

import numpy as np
import pytest
import torch

import hyperreel_b200 as hb
from hyperreel_b200 import lib as L
from oracle.hyperreel_oracle import HyperReelOracle
from tests.cases import build_case
from tests.golden.make_golden_grads import N_RAYS, probe_indices, target_for
from tests.test_parity_gpu import make_render

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SUPPORTED = ["technicolor_trained", "donerf_s16", "neural3d_trained"]


def _loss(rgb, n):
    return ((rgb - target_for(n).to(rgb.device)) ** 2).mean()


@pytest.mark.parametrize("name", SUPPORTED)
def test_parameter_gradients_match_reference_autograd(name):
    g = np.load(os.path.join(GOLDEN, f"grads_{name}.npz"))
    case = build_case(name)
    rays = case.rays[:N_RAYS].clone().cuda()
    render = make_render(case).cuda()
    rgb = render.model.render_differentiable(rays, clamp_output=True)  # eval-mode forward, like the golden
    loss = _loss(rgb, rays.shape[0])
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-5
    named = dict(render.named_parameters())
    keys = [k[len("norm/"):] for k in g.files if k.startswith("norm/")]
    assert len(keys) >= 17
    for k in keys:
        assert k in named, k
        grad = named[k].grad
        assert grad is not None, k
        flat = grad.reshape(-1).cpu()
        scale = float(g[f"max/{k}"]) + 1e-12
        assert abs(float(flat.norm()) - float(g[f"norm/{k}"])) <= 2e-3 * float(g[f"norm/{k}"]) + 1e-9, k
        probe = flat[probe_indices(flat.numel())].numpy()
        assert np.abs(probe - g[f"probe/{k}"]).max() <= 1e-3 * scale + 1e-10, k


@pytest.mark.parametrize("name", ["technicolor_app", "donerf_app", "neural3d_app", "donerf_distance", "technicolor_zero_net"])
@pytest.mark.parametrize("white", [False, True])
def test_training_mode_gradients_match_the_oracle(name, white):
    """training_step semantics (no clamp, optional white background) on the appearance-sensitive cases; every parameter
    tensor compared entry by entry with the oracle's autograd, plus d loss / d (sample-net output)."""
    case = build_case(name, n=200)
    rays = case.rays.clone()
    orc = HyperReelOracle(case.model_cfg_plain, case.dataset, case.state_dict)
    rgb_h, leaves_h = orc.render_with_grad(rays, clamp=False, white_bg=white, heads_leaf=True)
    _loss(rgb_h, rays.shape[0]).backward()
    rgb_o, leaves = orc.render_with_grad(rays, clamp=False, white_bg=white)
    _loss(rgb_o, rays.shape[0]).backward()
    render = make_render(case).cuda()
    render.train()
    rgb, heads = render.model.render_differentiable(rays.cuda(), white_bg=white, return_heads=True)
    heads.retain_grad()
    assert float((rgb.detach().cpu() - rgb_o.detach()).abs().max()) <= 2e-5
    _loss(rgb, rays.shape[0]).backward()
    ref_h = leaves_h["_mlp_out"].grad
    err_h = float((heads.grad.cpu() - ref_h).abs().max())
    # A `zero` net puts every sample exactly on its base plane, the last one on the aabb's max face (z = +1): there the
    # bilinear interpolation has a kink (grid_sample differentiates towards the zero padding, the kernel towards the interior
    # texel), and no parameter sits behind these heads anyway -- the table gradients below are what training uses.
    if case.sig.cfg.mlp_mode != L.MLP_ZERO:
        assert err_h <= 2e-3 * float(ref_h.abs().max()), f"d loss / d heads: {err_h} vs {float(ref_h.abs().max())}"
    for k, p in render.named_parameters():
        if k not in leaves or leaves[k].grad is None:
            continue
        ref = leaves[k].grad
        scale = float(ref.abs().max()) + 1e-12
        assert p.grad is not None, k
        err = float((p.grad.cpu() - ref).abs().max())
        assert err <= 2e-3 * scale, f"{k}: {err} vs scale {scale}"


def test_unsupported_pipelines_refuse_to_train():
    case = build_case("immersive_sphere_new", n=32)
    render = make_render(case).cuda()
    rgb = render.model.render_differentiable(case.rays.cuda())
    with pytest.raises(RuntimeError):
        rgb.sum().backward()


def test_optimizer_step_changes_the_next_render():
    """One Adam step on every parameter group: the library notices the new values (version counters) and re-packs."""
    case = build_case("technicolor_app", n=512)
    render = make_render(case).cuda()
    render.train()
    rays = case.rays.cuda()
    opt = torch.optim.Adam(render.parameters(), lr=3e-4)
    target = torch.full((rays.shape[0], 3), 0.25, device="cuda")
    losses = []
    for _ in range(8):
        opt.zero_grad(set_to_none=True)
        loss = ((render.model.render_differentiable(rays, white_bg=False) - target) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]
    render.eval()
    with torch.no_grad():
        a = render(rays)["rgb"]
    ref = HyperReelOracle(case.model_cfg_plain, case.dataset, {k: v.detach().cpu() for k, v in render.state_dict().items()}).render(case.rays.clone())
    assert float((a.cpu() - ref).abs().max()) <= 1e-4


def test_system_training_step_mirrors_the_reference_loop():
    """INRSystem.training_step (nlf/__init__.py:634-709): image loss, one Adam per optimiser group, loss goes down on a
    fixed batch, and the updated model still renders what the oracle computes from the updated parameters."""
    case = build_case("donerf_app", n=2048)
    cfg = hb.to_cfg({"model": case.model_cfg, "training": {"ray_chunk": 700, "iters_per_epoch": 4000,
                                                          "optimizers": {"color": {"lr": 0.002}, "color_impl": {"lr": 0.001},
                                                                         "embedding_impl": {"lr": 0.0002}}},
                     "dataset": case.dataset})
    system = hb.INRSystem(cfg)
    system.load_state_dict(case.state_dict)
    system.cuda()
    assert [len(o.param_groups[0]["params"]) > 0 for o in system.configure_optimizers()] == [True, True, True]
    g = torch.Generator().manual_seed(0)
    batch = {"coords": case.rays.cuda(), "rgb": torch.rand(case.rays.shape[0], 3, generator=g).cuda(),
             "weight": torch.ones(case.rays.shape[0], 1).cuda()}
    losses = [float(system.training_step(batch)["train/loss"]) for _ in range(6)]
    assert losses[-1] < losses[0], losses
    system.eval()
    with torch.no_grad():
        a = system(case.rays.cuda())["rgb"].cpu()
    sd = {k[len("render_fn."):]: v.detach().cpu() for k, v in system.state_dict().items()}
    ref = HyperReelOracle(case.model_cfg_plain, case.dataset, sd).render(case.rays.clone())
    assert float((a - ref).abs().max()) <= 1e-4


def test_training_across_a_grid_upsampling_step_with_the_tensorf_regulariser():
    """The schedule half of the reference's loop: `set_train_iter` re-samples the tables at an `upsamp_list` iteration
    (tensorf_base.py:509-553,1151-1188), rebuilds the occupancy mask and shrinks the aabb at an `update_AlphaMask_list` iteration
    (:379-429,1190-1232), the optimisers restart on the new Parameter objects, the TensoRF regulariser
    (nlf/regularizers/tensorf.py:35-96) adds its L1 / TV terms, training continues on the fused kernels at the new grid, and the
    re-packed model renders what the oracle computes from the up-sampled parameters."""
    case = build_case("donerf_app", n=1024)
    mcfg = hb.to_cfg(hb.config.to_plain(case.model_cfg))
    mcfg.color.net.N_voxel_init, mcfg.color.net.N_voxel_final = 40 ** 3, 56 ** 3
    mcfg.color.net.upsamp_list, mcfg.color.net.lr_upsample_reset = [3], True
    mcfg.color.net.update_AlphaMask_list = [5]  # occupancy mask + aabb shrink two iterations after the re-sampling
    mcfg.color.net.alpha_mask_thre = 1e-6       # (seeded tables are far less dense than a trained scene)
    reg = {"type": "tensorf", "update_AlphaMask_list": [5], "lr_decay_target_ratio": 0.1, "n_iters": 30000,
           "L1_weight_initial": 8e-5, "L1_weight_rest": 4e-5, "TV_weight_density": 0.05, "TV_weight_app": 0.05}
    cfg = hb.to_cfg({"model": mcfg, "training": {"ray_chunk": 1 << 20, "iters_per_epoch": 4000,
                                                 "optimizers": {"color": {"lr": 0.002}, "color_impl": {"lr": 0.001},
                                                                "embedding_impl": {"lr": 0.0002}}},
                     "dataset": case.dataset, "regularizers": {"tensorf": reg}})
    system = hb.INRSystem(cfg)
    system.load_state_dict(case.state_dict)
    system.cuda()
    net = system.render_fn.model.color_model.net
    grid0 = net.gridSize.tolist()
    g = torch.Generator().manual_seed(0)
    batch = {"coords": case.rays.cuda(), "rgb": torch.rand(case.rays.shape[0], 3, generator=g).cuda()}
    losses = [float(system.training_step(batch, train_iter=i)["train/loss"]) for i in range(7)]
    grid1 = net.gridSize.tolist()
    assert grid1 != grid0 and all(b > a for a, b in zip(grid0, grid1)), (grid0, grid1)
    assert net.density_plane[0].shape[-1] == grid1[0] and net.density_line[0].shape[2] == grid1[2]
    assert net.alphaMask is not None and system.regularizers[0].L1_reg_weight == 4e-5  # the pruning step ran at iteration 5
    assert losses[-1] < losses[0] and losses[-1] < losses[3], losses  # keeps improving after the re-sampling at iteration 3
    system.eval()
    with torch.no_grad():
        a = system(case.rays.cuda())["rgb"].cpu()
    sd = {k[len("render_fn."):]: v.detach().cpu() for k, v in system.state_dict().items()}
    ref = HyperReelOracle(hb.config.to_plain(mcfg), case.dataset, sd).render(case.rays.clone())
    assert float((a - ref).abs().max()) <= 1e-4
