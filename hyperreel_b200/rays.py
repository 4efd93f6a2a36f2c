"""Synthetic ray batches in the layouts the reference's datasets emit.

``forward_facing``: datasets/random.py:462-497 (RandomRayLightfieldDataset.get_random_rays) -- origins on the
st-plane z=-1, directions towards the uv-plane z=0, optional (camera_id, time) channels with time quantised
to frame centres (datasets/technicolor.py:122).  ``inward_360``: datasets/random.py:111-125 style -- Gaussian
origins, normalised Gaussian directions.  All generation is on the CPU with an explicit generator so the same
seed gives the same rays everywhere.
"""
from __future__ import annotations

import torch


def forward_facing(n: int, seed: int = 1, video: bool = True, num_frames: int = 50, pos_range: float = 0.25,
                   dir_range: float = 0.5) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    st = (torch.rand(n, 2, generator=g) * 2 - 1) * pos_range
    uv = (torch.rand(n, 2, generator=g) * 2 - 1) * dir_range
    o = torch.cat([st, -torch.ones(n, 1)], -1)
    d = torch.nn.functional.normalize(torch.cat([uv - st, torch.ones(n, 1)], -1), p=2.0, dim=-1)
    if not video:
        return torch.cat([o, d], -1).contiguous()
    t = torch.rand(n, 1, generator=g)
    if num_frames > 1:
        t = torch.round(t * (num_frames - 1)) / (num_frames - 1)
    return torch.cat([o, d, torch.zeros(n, 1), t], -1).contiguous()


def inward_360(n: int, seed: int = 1, pos_std: float = 0.3) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n, 3, generator=g) * pos_std
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), p=2.0, dim=-1)
    return torch.cat([o, d], -1).contiguous()


def for_signature(sig, n: int, seed: int = 1) -> torch.Tensor:
    """Rays of the right layout for a recognised pipeline."""
    from . import lib as L

    if sig.cfg.isect_type == L.ISECT_SPHERE:
        r = inward_360(n, seed)
        if sig.c_in == 8:
            g = torch.Generator().manual_seed(seed + 1000)
            r = torch.cat([r, torch.zeros(n, 1), torch.rand(n, 1, generator=g)], -1)
        return r
    r = forward_facing(n, seed, video=(sig.c_in == 8), num_frames=max(int(sig.cfg.num_frames), 1))
    if sig.cfg.n_color_views > 0:  # per-camera colour transform (point.py:594-605): spread the rays over the cameras
        r[:, 6] = (torch.arange(n) % int(sig.cfg.n_color_views)).float()
    return r
