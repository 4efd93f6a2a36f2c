"""hyperreel_b200 -- B200-native drop-in for HyperReel's per-ray rendering hot path.

Operator surface (same names as the reference's ``nlf`` package):
``render_fn_dict`` / ``RenderLightfield`` / ``render_chunked`` (rendering.py), ``model_dict`` /
``LightfieldModel`` (models.py), ``INRSystem`` (system.py).  All compute is in
``libhyperreel_b200.so`` (csrc/, sm_100a CUDA behind the C-ABI of include/hyperreel_b200.h).
"""
from . import camera, configs, rays  # noqa: F401
from .camera import Camera, generate_rays  # noqa: F401
from .config import Cfg, epochs_to_iters, load_model_yaml, to_cfg  # noqa: F401
from .models import LightfieldModel, model_dict  # noqa: F401
from .rendering import RenderLightfield, render_chunked, render_fn_dict  # noqa: F401
from .signature import Signature, UnsupportedPipeline, lower  # noqa: F401
from .system import INRSystem  # noqa: F401

__all__ = ["camera", "Camera", "generate_rays", "configs", "rays", "Cfg", "to_cfg", "load_model_yaml", "epochs_to_iters", "LightfieldModel", "model_dict",
           "RenderLightfield", "render_chunked", "render_fn_dict", "Signature", "UnsupportedPipeline", "lower",
           "INRSystem"]
