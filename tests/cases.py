"""Seeded parity cases shared by the golden generator, the oracle tests and the GPU parity tests.

Every case is (built-in model config in the reference schema, dataset facts, overrides, ray count, parameter
seed, density gain).  ``density_gain`` > 1 gives the "trained-like" variant in which transmittance saturates
along the ray, so all samples -- not only the last one -- contribute to the pixel.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Dict

import torch

from hyperreel_b200 import configs, rays as rays_mod
from hyperreel_b200.config import to_plain
from hyperreel_b200.signature import Signature, lower
from hyperreel_b200.state import seeded_state_dict

CASES: Dict[str, dict] = {
    # name: builtin, overrides, rays, parameter seed, density gain
    "technicolor_trained": dict(builtin="technicolor_z_plane", over=dict(n_voxels=48 ** 3), n=256, seed=0, gain=30.0),
    "technicolor_init": dict(builtin="technicolor_z_plane", over=dict(n_voxels=32 ** 3), n=128, seed=3, gain=1.0),
    "technicolor_k50": dict(builtin="technicolor_z_plane", over=dict(n_voxels=32 ** 3, num_keyframes=50), n=128, seed=4, gain=30.0),
    "neural3d_trained": dict(builtin="neural_3d_z_plane", over=dict(n_voxels=40 ** 3), n=160, seed=1, gain=30.0),
    "donerf_trained": dict(builtin="donerf_sphere", over=dict(n_voxels=40 ** 3), n=256, seed=2, gain=30.0),
    "donerf_s16": dict(builtin="donerf_sphere", over=dict(n_voxels=32 ** 3, z_channels=16), n=128, seed=5, gain=30.0),
    "plumbing_4096x4": dict(builtin="donerf_sphere", over=dict(n_voxels=64 ** 3, z_channels=4), n=4096, seed=6, gain=30.0),
    "shiny_tiny": dict(builtin="shiny_z_plane_tiny", over=dict(n_voxels=32 ** 3), n=256, seed=7, gain=30.0),
    # SURVEY 8 f3 families (config variants of the built-ins, hyperreel_b200/configs.py:_apply_variant)
    "technicolor_basic_pe": dict(builtin="technicolor_z_plane", over=dict(n_voxels=32 ** 3, variant="basic_pe"), n=128, seed=8, gain=30.0),
    "technicolor_bbox": dict(builtin="technicolor_z_plane", over=dict(n_voxels=32 ** 3, variant="bbox"), n=128, seed=9, gain=30.0),
    "technicolor_z_depth": dict(builtin="technicolor_z_plane", over=dict(n_voxels=32 ** 3, variant="z_depth"), n=128, seed=10, gain=30.0),
    "donerf_cylinder": dict(builtin="donerf_sphere", over=dict(n_voxels=32 ** 3, variant=["cylinder", "outward_facing"]), n=192, seed=11, gain=30.0),
    "technicolor_global_color": dict(builtin="technicolor_z_plane", over=dict(n_voxels=32 ** 3, variant="global_color"), n=128, seed=12, gain=30.0),
    "technicolor_both_color": dict(builtin="technicolor_z_plane", over=dict(n_voxels=32 ** 3, variant="both_color"), n=128, seed=13, gain=30.0),
    "shiny_scale_mask": dict(builtin="shiny_z_plane_tiny", over=dict(n_voxels=32 ** 3, variant="scale_mask"), n=192, seed=15, gain=30.0),
    "technicolor_z_scale": dict(builtin="technicolor_z_plane", over=dict(n_voxels=32 ** 3, variant="z_scale"), n=128, seed=16, gain=30.0),
    "immersive_sphere_new": dict(builtin="neural_3d_z_plane", over=dict(n_voxels=32 ** 3, z_channels=32, variant=["sphere_new", "outward_facing"]), n=160, seed=17, gain=30.0),
    "donerf_sphere_new": dict(builtin="donerf_sphere", over=dict(n_voxels=32 ** 3, variant="sphere_new"), n=192, seed=18, gain=30.0),
    # sample-net shapes beyond the headline one: encoded input wider than 32 features (two input chunks on the tensor-core
    # path), with and without the time group
    "donerf_wide_pe": dict(builtin="donerf_sphere", over=dict(n_voxels=32 ** 3, variant="wide_pe"), n=192, seed=19, gain=30.0),
    "neural3d_wide_pe": dict(builtin="neural_3d_z_plane", over=dict(n_voxels=32 ** 3, z_channels=32, variant="wide_pe"), n=160, seed=20, gain=30.0),
    # appearance-sensitive family: density high enough that sum(w) -> 1 and appearance tables at O(1) feature scale, so that
    # an error in the appearance gather shows in RGB far above the 1e-4 gate (tests/test_parity_bites.py quantifies it)
    "technicolor_app": dict(builtin="technicolor_z_plane", over=dict(n_voxels=48 ** 3), n=256, seed=21, gain=600.0, app_gain=6.0),
    "neural3d_app": dict(builtin="neural_3d_z_plane", over=dict(n_voxels=40 ** 3), n=160, seed=22, gain=100.0, app_gain=6.0),
    "donerf_app": dict(builtin="donerf_sphere", over=dict(n_voxels=40 ** 3), n=256, seed=23, gain=100.0, app_gain=10.0),
    "technicolor_zero_net": dict(builtin="technicolor_z_plane", over=dict(n_voxels=32 ** 3, variant="zero_net"), n=128, seed=24, gain=600.0, app_gain=6.0),
    "donerf_distance": dict(builtin="donerf_sphere", over=dict(n_voxels=32 ** 3, variant="distance"), n=192, seed=25, gain=100.0, app_gain=6.0),
    # two rays per warp (S <= 16) with several VM groups and SH shading: per-ray folded appearance matrices for both rays of a warp
    "neural3d_s16": dict(builtin="neural_3d_z_plane", over=dict(n_voxels=32 ** 3, z_channels=16), n=161, seed=26, gain=100.0, app_gain=6.0),
    "technicolor_s8": dict(builtin="technicolor_z_plane", over=dict(n_voxels=32 ** 3, z_channels=8), n=131, seed=27, gain=600.0, app_gain=6.0),
    "immersive_sphere_like": dict(builtin="neural_3d_z_plane", over=dict(n_voxels=32 ** 3, z_channels=32, variant=["sphere", "outward_facing"]), n=160, seed=14, gain=30.0),
}


@dataclass
class Case:
    name: str
    model_cfg: object
    model_cfg_plain: dict
    dataset: dict
    sig: Signature
    rays: torch.Tensor
    state_dict: Dict[str, torch.Tensor]
    n_samples: int


# render_kwargs of the extra-field fixtures (tensorf_dynamic.py:808-837): composited, raw per-sample and pred-weighted
# outputs, including a head that only reaches the colour net because it is named here (point.py:236-244)
FIELD_KWARGS = {"fields": ["render_weights", "distances", "points", "sigma", "viewdirs", "point_offset"],
                "no_over_fields": ["sigma"], "pred_weights_fields": ["viewdirs"]}


def state_hash(sd: Dict[str, torch.Tensor]) -> str:
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_case(name: str, n: int = None) -> Case:
    spec = CASES[name]
    cfg, ds = configs.get(spec["builtin"], **spec["over"])
    sig = lower(cfg, ds)
    sd = seeded_state_dict(sig, seed=spec["seed"], density_gain=spec["gain"], app_gain=spec.get("app_gain", 1.0))
    r = rays_mod.for_signature(sig, n or spec["n"], seed=100 + spec["seed"])
    return Case(name=name, model_cfg=cfg, model_cfg_plain=to_plain(cfg), dataset=ds, sig=sig, rays=r, state_dict=sd,
                n_samples=sig.n_samples)
