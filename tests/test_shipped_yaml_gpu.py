"""CUDA path vs the unmodified reference for every model YAML the reference ships and the fused path accepts
(tests/golden/shipped/*.npz, see tests/test_shipped_yaml_golden.py for the fixture format and the CPU half).

These fixtures were added after round 1's GPU budget was spent, so they have not run on hardware yet: the tests are
non-strict xfail (XPASS = parity holds, XFAIL = a combination that still needs work -- [8,8,8] components, S = 48 and
encoded inputs wider than 32 are exercised here for the first time).  Round 2: read the outcome, fix, drop the marker."""
import os

import pytest
import torch

import hyperreel_b200 as hb
from tests.test_parity_gpu import RGB_TOL
from tests.test_shipped_yaml_golden import SHIPPED, load_fixture

pytestmark = pytest.mark.gpu
PENDING = pytest.mark.xfail(strict=False, reason="first hardware run pending (added after the round-1 GPU budget was spent)")


def _render(cfg, ds, sd, rays, mode):
    model = hb.LightfieldModel(cfg, dataset=ds, mlp_mode=mode)
    render = hb.RenderLightfield(model, None, cfg.render, net_chunk=1 << 20)
    _, unexpected = render.load_state_dict(sd, strict=False)
    assert not unexpected
    render.eval()
    return render(rays.cuda())["rgb"].cpu()


@PENDING
@pytest.mark.parametrize("path", SHIPPED, ids=[os.path.basename(p)[:-4] for p in SHIPPED])
def test_shipped_yaml_fp32_path_matches_reference(path):
    plain, cfg, ds, sig, sd, rays, rgb = load_fixture(path)
    out = _render(cfg, ds, sd, rays, "fp32")
    assert float((out - rgb).abs().max()) <= RGB_TOL


@PENDING
@pytest.mark.parametrize("path", SHIPPED, ids=[os.path.basename(p)[:-4] for p in SHIPPED])
def test_shipped_yaml_tensor_core_path_matches_reference(path):
    plain, cfg, ds, sig, sd, rays, rgb = load_fixture(path)
    if sig.cfg.mlp_width != 256 or sig.cfg.mlp_in > 32:
        pytest.skip("tensor-core sample net needs width 256 and an encoded input of at most 32 features")
    out = _render(cfg, ds, sd, rays, "bf16x3")
    assert float((out - rgb).abs().max()) <= RGB_TOL
