"""Golden GRADIENTS for the backward pass of the path (not built yet): d loss / d parameter from the
reference's own autograd (unmodified modules through oracle/ref_shim.py, CPU, eval-mode forward), loss = mean((rgb - target)^2)
with a seeded target.  Per parameter: L2 norm, max |g| and 64 probe entries (seeded indices).  The gradient oracle
(HyperReelOracle.render_with_grad) is checked against these where the reference is absent.

    python tests/golden/make_golden_grads.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from tests.cases import build_case  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
GRAD_CASES = ["technicolor_trained", "donerf_s16", "neural3d_trained", "immersive_sphere_new"]
N_RAYS = 96


def target_for(n):
    g = torch.Generator().manual_seed(1234)
    return torch.rand(n, 3, generator=g)


def probe_indices(numel, k=64):
    g = torch.Generator().manual_seed(numel % 100003)
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def main():
    ref_shim.install()
    for name in GRAD_CASES:
        case = build_case(name)
        rays = case.rays[:N_RAYS].clone()
        ref = ref_shim.build_reference(case.model_cfg_plain, case.dataset)
        ref.load_state_dict(case.state_dict, strict=False)
        ref.eval()
        for p in ref.parameters():
            p.requires_grad_(True)
        out = ref(rays.clone())["rgb"].reshape(-1, 3)
        loss = ((out - target_for(rays.shape[0])) ** 2).mean()
        loss.backward()
        rec = {"loss": np.array(float(loss))}
        for k, p in ref.named_parameters():
            if p.grad is None:
                continue
            g = p.grad.detach().reshape(-1)
            rec[f"norm/{k}"] = np.array(float(g.norm()))
            rec[f"max/{k}"] = np.array(float(g.abs().max()))
            rec[f"probe/{k}"] = g[probe_indices(g.numel())].numpy()
        np.savez_compressed(os.path.join(OUT, f"grads_{name}.npz"), **rec)
        print(name, "loss", float(loss), "params with grad", sum(1 for k in rec if k.startswith("norm/")))


if __name__ == "__main__":
    main()
