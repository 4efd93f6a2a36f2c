"""Gradient oracle (HyperReelOracle.render_with_grad) against gradients from the reference's own autograd
(tests/golden/grads_*.npz, tests/golden/make_golden_grads.py).  Test infrastructure for the backward pass of the path
(SURVEY.md 8 f1); no product code is exercised here."""
import os

import numpy as np
import pytest
import torch

from oracle.hyperreel_oracle import HyperReelOracle
from tests.cases import build_case
from tests.golden.make_golden_grads import GRAD_CASES, N_RAYS, probe_indices, target_for

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", GRAD_CASES)
def test_oracle_gradients_match_reference_autograd(name):
    g = np.load(os.path.join(GOLDEN, f"grads_{name}.npz"))
    case = build_case(name)
    rays = case.rays[:N_RAYS].clone()
    orc = HyperReelOracle(case.model_cfg_plain, case.dataset, case.state_dict)
    rgb, leaves = orc.render_with_grad(rays)
    loss = ((rgb - target_for(rays.shape[0])) ** 2).mean()
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-6
    keys = [k[len("norm/"):] for k in g.files if k.startswith("norm/")]
    assert len(keys) >= 17
    for k in keys:
        grad = leaves[k].grad
        assert grad is not None, k
        flat = grad.reshape(-1)
        scale = float(g[f"max/{k}"]) + 1e-12
        assert abs(float(flat.norm()) - float(g[f"norm/{k}"])) <= 1e-4 * float(g[f"norm/{k}"]) + 1e-9, k
        probe = flat[probe_indices(flat.numel())].detach().numpy()
        assert np.abs(probe - g[f"probe/{k}"]).max() <= 2e-5 * scale + 1e-10, k
