"""``render_fn_dict['lightfield']`` and ``render_chunked`` of the drop-in.

Mirrors nlf/rendering.py:59-150: ``RenderLightfield(model, subdivision, cfg.model.render, net_chunk=int)`` is an
``nn.Module`` whose ``forward(rays[N,C], **render_kwargs)`` returns a dict of ``[N, .]`` tensors and which
exposes ``.embed`` / ``.forward_multiple`` and ``.model``; ``render_chunked`` is the chunk loop every caller
of the reference goes through (``INRSystem.run_chunked``, nlf/__init__.py:491-502).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict

import torch
from torch import nn


class Render(nn.Module):
    def __init__(self, model, subdivision, cfg, **kwargs):
        super().__init__()
        self.net_chunk = kwargs.get("net_chunk", 32768)

    @staticmethod
    def _flat(out: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return {k: v.reshape(-1, v.shape[-1]) for k, v in out.items()}


class RenderLightfield(Render):
    def __init__(self, model, subdivision, cfg, *args, **kwargs):
        super().__init__(model, subdivision, cfg, **kwargs)
        if subdivision is not None:
            raise NotImplementedError("subdivision is not on the fused path")
        self.model = model

    def forward(self, rays, **render_kwargs):
        return self._flat(self.model(rays.reshape(-1, rays.shape[-1]), render_kwargs))

    def embed(self, rays, **render_kwargs):
        return self._flat(self.model.embed(rays.reshape(-1, rays.shape[-1]), render_kwargs))

    def forward_multiple(self, rays, **render_kwargs):
        return self.forward(rays, **render_kwargs)


render_fn_dict = {"lightfield": RenderLightfield}


def render_chunked(rays, render_fn, render_kwargs, chunk):
    """Chunk loop + per-key concatenation (nlf/rendering.py:100-150).  Output is identical for any
    ``chunk`` because rays are independent; the fused kernels make large chunks cheap, so callers should
    pass ``chunk >= 65536`` (the reference default of 16 384 only bounded PyTorch's activation memory)."""
    B = rays.shape[0]
    chunk = int(chunk) if chunk and chunk > 0 else max(B, 1)
    if B <= chunk:
        return dict(render_fn(rays, **render_kwargs))
    results = defaultdict(list)
    for i in range(0, B, chunk):
        for k, v in render_fn(rays[i:i + chunk], **render_kwargs).items():
            results[k].append(v)
    return {k: torch.cat(v, 0) for k, v in results.items()}
