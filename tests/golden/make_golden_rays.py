"""Golden vectors for the camera -> rays step: outputs of the reference's own utils/ray_utils.py functions
(get_ray_directions_K, get_rays, get_ndc_rays_fx_fy) composed as in datasets/base.py:485-518, on CPU through the shim.

    python tests/golden/make_golden_rays.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from tests.cases_rays import RAY_CASES  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref_shim.install()
    from utils.ray_utils import get_ndc_rays_fx_fy, get_ray_directions_K, get_rays

    for name, c in RAY_CASES.items():
        K = torch.FloatTensor(c["K"])
        c2w = torch.FloatTensor(c["pose"])[:3, :4]
        directions = get_ray_directions_K(c["H"], c["W"], K, centered_pixels=True, device="cpu")
        rays_o, rays_d = get_rays(directions, c2w)
        rays = torch.cat([rays_o, rays_d], dim=-1)
        if c["use_ndc"]:
            rays = get_ndc_rays_fx_fy(c["H"], c["W"], K[0, 0], K[1, 1], c["near"], rays)
        rays = torch.cat([rays, torch.ones_like(rays[..., :1]) * c["cam_idx"]], dim=-1)
        rays = torch.cat([rays, torch.ones_like(rays[..., :1]) * c["time"]], dim=-1)
        np.savez_compressed(os.path.join(OUT, f"rays_{name}.npz"), rays=rays.numpy())
        print(name, tuple(rays.shape), float(rays.abs().max()))


if __name__ == "__main__":
    main()
