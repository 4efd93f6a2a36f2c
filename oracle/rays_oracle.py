"""CPU oracle for the steps either side of the hot path.  TEST INFRASTRUCTURE.

Restates, in plain torch CPU arithmetic, the reference's camera-to-ray generation
(utils/ray_utils.py:98-164 as driven by datasets/base.py:485-518) and its 8-bit packing
(utils/__init__.py:47).  Pinned against the reference functions themselves through oracle/ref_shim.py
(tests/test_oracle_vs_reference.py) and against tests/golden/rays_*.npz everywhere else.
"""
from __future__ import annotations

import numpy as np
import torch


def pixel_grid(H: int, W: int) -> torch.Tensor:
    """kornia.create_meshgrid(H, W, normalized_coordinates=False)[0]: [H, W, 2] with (x, y)."""
    xs = torch.linspace(0, W - 1, W)
    ys = torch.linspace(0, H - 1, H)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([gx, gy], -1)


def ray_directions(H, W, K, centered_pixels=False, flipped=False):
    """get_ray_directions_K -> get_ray_directions_from_pixels_K (ray_utils.py:98-119)."""
    grid = pixel_grid(H, W)
    i, j = grid.unbind(-1)
    off = 0.5 if centered_pixels else 0.0
    dy = (j - K[1, 2] + off) / K[1, 1]
    return torch.stack([(i - K[0, 2] + off) / K[0, 0], dy if flipped else -dy, -torch.ones_like(i)], -1)


def get_rays(directions, c2w, normalize=True):
    """ray_utils.py:121-135."""
    rays_d = directions @ c2w[:, :3].T
    if normalize:
        rays_d = rays_d / torch.clamp(torch.sqrt((rays_d * rays_d).sum(-1, keepdim=True)), min=1e-12)
    rays_o = c2w[:, 3].expand(rays_d.shape)
    return rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)


def ndc_rays(H, W, fx, fy, near, rays):
    """get_ndc_rays_fx_fy (ray_utils.py:137-164)."""
    rays_o, rays_d = rays[..., 0:3], rays[..., 3:6]
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    ox_oz = rays_o[..., 0] / rays_o[..., 2]
    oy_oz = rays_o[..., 1] / rays_o[..., 2]
    o0 = -1. / (W / (2. * fx)) * ox_oz
    o1 = -1. / (H / (2. * fy)) * oy_oz
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * fx)) * (rays_d[..., 0] / rays_d[..., 2] - ox_oz)
    d1 = -1. / (H / (2. * fy)) * (rays_d[..., 1] / rays_d[..., 2] - oy_oz)
    d2 = 1 - o2
    return torch.cat([torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)], -1)


def coords_from_camera(pose, K, W, H, time=0.0, cam_idx=0.0, use_ndc=False, near=1.0, centered_pixels=True, c_in=8):
    """get_coords_from_camera (datasets/base.py:485-518)."""
    K = torch.as_tensor(K, dtype=torch.float32)
    c2w = torch.as_tensor(pose, dtype=torch.float32)[:3, :4]
    d = ray_directions(H, W, K, centered_pixels=centered_pixels)
    o, d = get_rays(d, c2w)
    rays = torch.cat([o, d], -1)
    if use_ndc:
        rays = ndc_rays(H, W, K[0, 0], K[1, 1], near, rays)
    if c_in == 8:
        rays = torch.cat([rays, torch.ones_like(rays[..., :1]) * cam_idx, torch.ones_like(rays[..., :1]) * time], -1)
    return rays


def to8b(x: np.ndarray) -> np.ndarray:
    """utils/__init__.py:47."""
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)
