"""The model YAMLs the reference ships and the fused path accepts (tests/golden/shipped/*.npz, generated from the unmodified
reference by tests/golden/make_golden_shipped.py): the configuration travels as JSON, parameters are re-seeded here.

CPU part (this file): the oracle reproduces the reference's rgb for every one of them, and every one still lowers.
The matching GPU check (CUDA path vs these fixtures) is listed under DESIGN.md section 8 "next"."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import hyperreel_b200 as hb
from hyperreel_b200.state import seeded_state_dict
from oracle.hyperreel_oracle import HyperReelOracle

SHIPPED = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shipped", "*.npz")))
PARAM_SEED = 3


def load_fixture(path):
    g = np.load(path)
    plain = json.loads(str(g["config_json"]))
    ds = json.loads(str(g["dataset_json"]))
    cfg = hb.to_cfg(plain)
    sig = hb.lower(cfg, ds)
    sd = seeded_state_dict(sig, seed=PARAM_SEED, density_gain=30.0)
    return plain, cfg, ds, sig, sd, torch.from_numpy(g["rays"]), torch.from_numpy(g["rgb"])


def test_fixture_set_is_complete():
    assert len(SHIPPED) == 45


@pytest.mark.parametrize("path", SHIPPED, ids=[os.path.basename(p)[:-4] for p in SHIPPED])
def test_oracle_reproduces_reference_rgb_for_shipped_yaml(path):
    plain, cfg, ds, sig, sd, rays, rgb = load_fixture(path)
    out = HyperReelOracle(plain, ds, sd).render(rays.clone())
    assert float((out - rgb).abs().max()) <= 2e-6
