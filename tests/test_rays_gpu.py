"""The steps either side of the path on the GPU: device ray generation (f2) and fused 8-bit packing (f4)."""
import os

import numpy as np
import pytest
import torch

import hyperreel_b200 as hb
from oracle.rays_oracle import coords_from_camera, to8b
from tests.cases import build_case
from tests.cases_rays import RAY_CASES
from tests.test_parity_gpu import make_render

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _camera(c):
    return hb.Camera(pose=c["pose"], K=c["K"], width=c["W"], height=c["H"], time=c["time"], cam_idx=c["cam_idx"],
                     use_ndc=c["use_ndc"], ndc_near=c["near"])


@pytest.mark.parametrize("name", list(RAY_CASES))
def test_device_rays_match_reference_golden(name):
    c = RAY_CASES[name]
    g = np.load(os.path.join(GOLDEN, f"rays_{name}.npz"))["rays"]
    rays = hb.generate_rays(_camera(c), c_in=8).cpu().numpy()
    assert rays.shape == g.shape
    assert np.abs(rays - g).max() <= 4e-6 * max(1.0, np.abs(g).max())
    rays6 = hb.generate_rays(_camera(c), c_in=6).cpu().numpy()
    assert np.array_equal(rays6, rays[:, :6])


def test_device_rays_pixel_subrange_and_errors():
    c = RAY_CASES["world_50x37"]
    full = hb.generate_rays(_camera(c), c_in=8)
    part = hb.generate_rays(_camera(c), c_in=8, first_pixel=123, n_pixels=777)
    assert torch.equal(part, full[123:900])
    with pytest.raises(RuntimeError):
        hb.generate_rays(_camera(c), c_in=8, first_pixel=1800, n_pixels=100)  # beyond the image
    with pytest.raises(RuntimeError):
        hb.generate_rays(_camera(c), c_in=7)


def test_to8b_epilogue_is_exactly_to8b_of_the_float_path():
    case = build_case("technicolor_trained", n=3000)
    render = make_render(case, mlp_mode="bf16x3")
    rays = case.rays.cuda()
    f = render(rays)["rgb"].cpu().numpy()
    u = render.model.render_to8b(rays).cpu().numpy()
    assert u.dtype == np.uint8 and u.shape == (3000, 3)
    assert np.array_equal(u, to8b(f))  # integer work: bit exact
    assert int(u.max()) > 0


def test_whole_frame_from_camera_matches_the_separate_steps():
    """hr_render_frame_to8b_host == generate_rays -> render -> to8b, and agrees with the CPU oracles end to end."""
    from oracle.hyperreel_oracle import HyperReelOracle
    case = build_case("technicolor_trained", n=8)
    render = make_render(case)
    cam = hb.Camera(pose=RAY_CASES["ndc_73x41"]["pose"], K=RAY_CASES["ndc_73x41"]["K"], width=73, height=41, time=0.25,
                    cam_idx=0.0, use_ndc=True, ndc_near=0.6)
    img = render.model.render_frame_to8b(cam, chunk=1000)  # 3 chunks
    assert img.shape == (41, 73, 3) and img.dtype == torch.uint8
    rays = hb.generate_rays(cam, c_in=8)
    sep = render.model.render_to8b(rays).cpu().reshape(41, 73, 3)
    assert torch.equal(img, sep)
    ref_rays = coords_from_camera(cam.pose, cam.K, 73, 41, 0.25, 0.0, True, 0.6)
    ref = HyperReelOracle(case.model_cfg_plain, case.dataset, case.state_dict).render(ref_rays)
    diff = np.abs(img.numpy().astype(np.int32).reshape(-1, 3) - to8b(ref.numpy()).astype(np.int32))
    assert diff.max() <= 1  # a 1e-4 float difference can move a value across one 8-bit boundary
