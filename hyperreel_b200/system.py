"""Render-path surface of the reference's ``INRSystem`` (nlf/__init__.py:278-502).

Only what sits on the hot path is kept: construction from the full config (``cfg.model``, ``cfg.training``,
``cfg.dataset``), ``forward`` / ``render`` / ``run_chunked`` with the reference's chunk selection, and
``load_state_dict`` with the reference's grid-size fix-up (nlf/__init__.py:433-479), including Lightning
checkpoints whose keys carry the ``render_fn.`` prefix.  Training, optimisers, regularisers, visualisers
and datasets are out of scope.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from .config import Cfg, epochs_to_iters, to_cfg
from .models import model_dict
from .rendering import render_chunked, render_fn_dict


class INRSystem(nn.Module):
    def __init__(self, cfg, dm=None, dataset: Optional[dict] = None, mlp_mode: str = "auto"):
        super().__init__()
        self.cfg = to_cfg(cfg)
        self.dm = dm
        training = self.cfg.get("training", Cfg())
        ipe = training.get("iters_per_epoch", None)
        if ipe is not None:
            epochs_to_iters(self.cfg, ipe)  # nlf/__init__.py:306-315
        if dataset is None and dm is None and "dataset" in self.cfg:
            d = self.cfg.dataset
            dataset = {k: d[k] for k in ("name", "collection", "num_keyframes", "num_frames", "near", "far", "depth_range") if k in d}
        model = model_dict[self.cfg.model.type](self.cfg.model, system=self if dm is not None else None,
                                                dataset=dataset, iters_per_epoch=ipe, mlp_mode=mlp_mode)
        self.rendering = False
        self.render_fn = render_fn_dict[self.cfg.model.render.type](
            model, None, self.cfg.model.render, net_chunk=training.get("net_chunk", 32768))
        self.eval()

    # ---- nlf/__init__.py:481-502
    def render(self, method_name, coords, **render_kwargs):
        return self.run_chunked(coords, getattr(self.render_fn, method_name), **render_kwargs)

    def forward(self, coords, **render_kwargs):
        return self.run_chunked(coords, self.render_fn, **render_kwargs)

    def run_chunked(self, coords, fn, **render_kwargs):
        training = self.cfg.get("training", Cfg())
        if self.rendering or render_kwargs.pop("rendering", False):
            chunk = training.get("render_ray_chunk", training.get("ray_chunk", 1 << 20))
        else:
            chunk = training.get("ray_chunk", 1 << 20)
        return render_chunked(coords, fn, render_kwargs, chunk=chunk)

    # ---- nlf/__init__.py:433-479
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = False):
        if "state_dict" in state_dict and isinstance(state_dict["state_dict"], dict):
            state_dict = state_dict["state_dict"]  # a Lightning .ckpt
        net = self.render_fn.model.color_model.net
        sd = {}
        for k, v in state_dict.items():
            k2 = k if k.startswith("render_fn.") else "render_fn." + k
            sd[k2] = v
            if k2.endswith("color_model.net.gridSize"):
                net.gridSize = torch.as_tensor(v, dtype=torch.long).cpu()
                net.init_svd_volume(net.gridSize[0], net.device)
        own = self.state_dict()
        for k in list(sd.keys()):
            if k in own and any(t in k for t in ("app_plane", "density_plane", "app_line", "density_line")):
                if sd[k].shape != own[k].shape:
                    sd[k] = sd[k].view(*own[k].shape)
        missing = super().load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
        net.update_stepSize(net.gridSize)
        self.render_fn.model.mark_dirty()
        return missing
