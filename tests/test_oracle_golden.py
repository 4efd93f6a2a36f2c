"""The oracle against the committed golden vectors (outputs of the reference itself, see
tests/golden/make_golden.py).  Runs everywhere, no GPU, no reference checkout needed."""
import os

import numpy as np
import pytest
import torch

from oracle.hyperreel_oracle import HyperReelOracle
from tests.cases import CASES, build_case, state_hash

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.mark.parametrize("name", list(CASES))
def test_seeded_parameters_match_fixture(name):
    case = build_case(name)
    g = load_golden(name)
    assert state_hash(case.state_dict) == str(g["state_sha256"]), "seeded parameters drifted: regenerate tests/golden"
    assert np.array_equal(case.rays.numpy(), g["rays"])


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("gather", ["explicit", "grid_sample"])
def test_oracle_matches_reference_golden(name, gather):
    case = build_case(name)
    g = load_golden(name)
    orc = HyperReelOracle(case.model_cfg_plain, case.dataset, case.state_dict, gather=gather)
    st = {}
    rgb = orc.render(case.rays.clone(), st)
    # fp32 restatement vs fp32 reference: only summation-order noise is allowed
    assert np.abs(rgb.numpy() - g["rgb"]).max() <= 2e-6
    assert np.abs(st["mlp_out"][:64].numpy() - g["mlp_out"]).max() <= 1e-5
    assert np.abs(st["distances"].numpy() - g["distances"]).max() <= 2e-6
    assert np.abs(st["points"].numpy() - g["points"]).max() <= 2e-6
    assert np.abs(st["weights"].numpy() - g["render_weights"]).max() <= 2e-6


@pytest.mark.parametrize("name", ["technicolor_trained", "donerf_trained"])
def test_fp64_oracle_bounds_reference_rounding(name):
    """The reference's own fp32 rounding noise (distance to an fp64 evaluation) is far below the 1e-4 gate."""
    case = build_case(name)
    g = load_golden(name)
    rgb64 = HyperReelOracle(case.model_cfg_plain, case.dataset, case.state_dict, dtype=torch.float64).render(case.rays.clone())
    assert np.abs(rgb64.float().numpy() - g["rgb"]).max() <= 1e-5
