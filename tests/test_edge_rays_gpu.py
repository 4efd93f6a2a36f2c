"""Crafted rays at the discontinuities of the path (SURVEY Appendix E "parity traps"), CUDA path vs the CPU oracle.

The oracle is pinned to the unmodified reference on the seeded cases (tests/test_oracle_vs_reference.py); here it checks
inputs the random ray generators never produce: rays parallel to the sample planes (the |d_z| < 1e-5 -> 1e12 guard,
utils/intersect_utils.py:135-142), times at the ends and at keyframe-snap boundaries (utils/flow_utils.py:18-31), origins far
outside / at the centre of the volume, un-normalised directions, rays that miss everything."""
import pytest
import torch

from oracle.hyperreel_oracle import HyperReelOracle
from tests.cases import build_case
from tests.test_parity_gpu import RGB_TOL, make_render

pytestmark = pytest.mark.gpu


def craft(case):
    r = case.rays.clone()
    n, c_in = r.shape
    assert n >= 96
    r[0:4, 5] = 0.0                      # exactly parallel to the z planes
    r[4:8, 5] = 5e-6                     # inside the guard band
    r[8:12, 5] = -5e-6
    r[12:16, 5] = 2e-5                   # just outside it
    r[16:24, 0:3] *= 40.0                # far origins (contraction tail / outside the aabb)
    r[24:28, 0:3] = 0.0                  # at the centre of the volume
    r[28:36, 3:6] *= 1e-3                # tiny, un-normalised directions
    r[36:40, 3:6] *= 250.0               # huge directions
    r[40:44, 3:6] = -r[40:44, 3:6]       # looking away
    r[44:48, 3] = 0.0
    r[44:48, 4] = 0.0                    # straight down the axis
    if c_in == 8:
        K, Fr = case.dataset["num_keyframes"], case.dataset["num_frames"]
        fac = K * (Fr - 1) / Fr
        r[48:52, 7] = 0.0
        r[52:56, 7] = 1.0
        for i in range(56, 80):          # around the keyframe rounding boundaries (k + 0.5) / fac
            k = (i - 56) % max(K - 1, 1)
            r[i, 7] = min(1.0, max(0.0, (k + 0.5) / fac + (1e-5 if i % 2 else -1e-5) + 1e-5 / fac))
        r[80:84, 7] = -0.25              # outside [0, 1] (clamped by get_base_time)
        r[84:88, 7] = 1.5
    return r.contiguous()


@pytest.mark.parametrize("name", ["technicolor_trained", "neural3d_trained", "donerf_trained", "immersive_sphere_new",
                                  "donerf_cylinder", "technicolor_bbox"])
@pytest.mark.parametrize("mode", ["fp32", "bf16x3"])
def test_crafted_rays_match_oracle(name, mode):
    case = build_case(name)
    rays = craft(case)
    render = make_render(case, mlp_mode=mode)
    got = render(rays.cuda())["rgb"].cpu()
    ref = HyperReelOracle(case.model_cfg_plain, case.dataset, case.state_dict).render(rays.clone())
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max(dim=1)[0]
    bad = torch.nonzero(err > RGB_TOL).flatten().tolist()
    assert not bad, f"{name} [{mode}]: rays {bad[:12]} differ, max {float(err.max()):.3e}"
