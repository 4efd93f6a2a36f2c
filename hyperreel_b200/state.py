"""Parameter containers with the reference's ``state_dict`` names, and reference-style initialisers.

The fused model keeps its parameters in ordinary ``nn.Parameter``s laid out exactly like the reference
modules, so a Lightning checkpoint of the reference loads with
``load_state_dict`` -- only the *storage* lives here; no torch op ever runs on them on the render path
(the native library packs them into its own device layout on upload).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from .signature import Signature

MAT_MODE = [[0, 1], [0, 2], [1, 2]]  # tensorf_base.py:231, tensorf_dynamic.py:47
VEC_MODE = [2, 1, 0]                 # tensorf_base.py:232; time planes pair axis c with time (tensorf_dynamic.py:48)


def n_to_reso(n_voxels: int, aabb: torch.Tensor) -> List[int]:
    """utils/tensorf_utils.py:65-68 (fp32 tensor arithmetic, truncating)."""
    xyz_min, xyz_max = aabb[0].float(), aabb[1].float()
    voxel_size = ((xyz_max - xyz_min).prod() / n_voxels).pow(1 / 3)
    return ((xyz_max - xyz_min) / voxel_size).long().tolist()


class _SampleNet(nn.Module):
    """Storage twin of BaseMLP (nlf/nets/mlp.py:127-154): layers.{i}.0.{weight,bias}, last layers.{L-1}.{...}."""

    def __init__(self, shapes: Sequence[tuple]):
        super().__init__()
        if len(shapes) == 0:  # ZeroMLP (nlf/nets/mlp.py:14-33): one unused Linear(1, 1) named `layer`
            self.layer = nn.Linear(1, 1)
            return
        self.layers = nn.ModuleList()
        for i, (fout, fin) in enumerate(shapes):
            lin = nn.Linear(fin, fout)  # PyTorch default init == reference (weight_init: none)
            self.layers.append(lin if i == len(shapes) - 1 else nn.Sequential(lin, nn.LeakyReLU(0.01)))


class _Prediction(nn.Module):
    def __init__(self, shapes):
        super().__init__()
        self.net = _SampleNet(shapes)


class _ColorTransform(nn.Module):
    """Storage twin of ColorTransformEmbedding (nlf/embedding/point.py:558-592): one 3x3 transform + shift per camera, zeros."""

    def __init__(self, views: int):
        super().__init__()
        self.color_embedding = nn.Parameter(torch.zeros(views, 12))


class _Embedding(nn.Module):
    """`embeddings` mirrors RayPointEmbedding's ModuleList (nlf/embedding/embedding.py:80-96): one entry per YAML key, in
    order; entries without parameters are empty modules (they add no state_dict keys).  Entry 0 is the ray_prediction; a
    cascaded pipeline has its point_prediction net at `net_index` (then `shapes` is that net's and `pre_shapes` the ray net's,
    [] for a `zero` net); a colour transform sits at `color_index`."""

    def __init__(self, shapes, color_views: int = 0, color_index: int = -1, net_index: int = 0, pre_shapes=None):
        super().__init__()
        entries = {0: _Prediction(shapes if net_index == 0 else (pre_shapes or []))}
        if net_index > 0:
            entries[net_index] = _Prediction(shapes)
        if color_views > 0 and color_index > 0:
            entries[color_index] = _ColorTransform(color_views)
        self.embeddings = nn.ModuleList([entries.get(i, nn.Module()) for i in range(max(entries) + 1)])


class _Tensorf(nn.Module):
    """Storage twin of TensorVMKeyframeTime / TensorVMNoSample parameters (tensorf_dynamic.py:106-244,
    tensorf_base.py:895-991)."""

    def __init__(self, sig: Signature, grid: Sequence[int]):
        super().__init__()
        c = sig.cfg
        self.dynamic = bool(c.dynamic)
        self.register_buffer("aabb", torch.tensor([[c.aabb[0], c.aabb[1], c.aabb[2]], [c.aabb[3], c.aabb[4], c.aabb[5]]]))
        self.register_buffer("gridSize", torch.tensor(list(grid), dtype=torch.long))
        self.n_sigma = [int(c.n_sigma[i]) for i in range(3)]
        self.n_app = [int(c.n_app[i]) for i in range(3)]
        self.K = int(c.num_keyframes)
        self.basis_mat = nn.Linear(sum(self.n_app), int(c.app_dim), bias=False)
        if self.dynamic:
            self.basis_mat_density = nn.Linear(sum(self.n_sigma), 1, bias=False)
        self.alphaMask = None
        self.device = "cuda"
        self.init_svd_volume(None, None)
        # up-sampling schedule (tensorf_base.py:150-197): log-spaced voxel counts (or per-axis grid sizes) between the initial and
        # the final grid, one step per entry of upsamp_list
        net = sig.model_cfg.color.net
        self.upsamp_list = [int(v) for v in (net.get("upsamp_list", []) or [])]
        self.lr_upsample_reset = bool(net.get("lr_upsample_reset", False))
        self.needs_opt_reset = False
        self.cur_iter = 0
        steps = len(self.upsamp_list) + 1
        import math
        if "grid_size" in net:
            self.use_grid_size_upsample = True
            self.N_voxel_list = [torch.round(torch.exp(torch.linspace(math.log(float(net.grid_size.start[i])), math.log(float(net.grid_size.end[i])),
                                                                        steps))).long().tolist()[1:] for i in range(3)]
        else:
            self.use_grid_size_upsample = False
            n0, n1 = float(net.get("N_voxel_init", 1)), float(net.get("N_voxel_final", net.get("N_voxel_init", 1)))
            self.N_voxel_list = torch.round(torch.exp(torch.linspace(math.log(n0), math.log(n1), steps))).long().tolist()[1:]

    def _shapes(self, comps, i, grid):
        a, b = MAT_MODE[i]
        v = VEC_MODE[i]
        plane = (1, comps[i], int(grid[b]), int(grid[a]))
        second = (1, comps[i], self.K, int(grid[v])) if self.dynamic else (1, comps[i], int(grid[v]), 1)
        return plane, second

    def init_svd_volume(self, res, device):
        """Re-create the tables at ``self.gridSize`` (reference: nlf/__init__.py:448-454 calls this before
        loading a checkpoint whose grids were up-sampled/shrunk)."""
        grid = self.gridSize.tolist()
        names = (("density_plane_space", "density_plane_time", "app_plane_space", "app_plane_time") if self.dynamic
                 else ("density_plane", "density_line", "app_plane", "app_line"))
        dp, d2, ap, a2 = [], [], [], []
        for i in range(3):
            ps, ss = self._shapes(self.n_sigma, i, grid)
            pa, sa = self._shapes(self.n_app, i, grid)
            # density: 1e-2 * U(0,1).clamp(1e-2, 1e8)  (fea2denseAct relu); appearance: 0.1 * N(0,1)
            dp.append(nn.Parameter(1e-2 * torch.rand(ps).clamp(1e-2, 1e8)))
            d2.append(nn.Parameter(1e-2 * torch.rand(ss).clamp(1e-2, 1e8)))
            ap.append(nn.Parameter(0.1 * torch.randn(pa)))
            a2.append(nn.Parameter(0.1 * torch.randn(sa)))
        setattr(self, names[0], nn.ParameterList(dp))
        setattr(self, names[1], nn.ParameterList(d2))
        setattr(self, names[2], nn.ParameterList(ap))
        setattr(self, names[3], nn.ParameterList(a2))
        self.struct_version = getattr(self, "struct_version", 0) + 1  # new Parameter objects: caches keyed on them are stale

    def update_stepSize(self, gridSize):
        self.gridSize = torch.as_tensor(gridSize, dtype=torch.long)

    # ---- grid up-sampling (tensorf_base.py:1151-1188 `up_sampling_VM` / `upsample_volume_grid`, tensorf_dynamic.py:394-441):
    # bilinear, align_corners=True re-sampling of every table, new Parameter objects (the optimiser must be rebuilt)
    @torch.no_grad()
    def upsample_volume_grid(self, res_target):
        import torch.nn.functional as F

        res = [int(v) for v in res_target]
        dp, d2, ap, a2 = self.tables()
        for planes, seconds, comps in ((ap, a2, self.n_app), (dp, d2, self.n_sigma)):
            for i in range(3):
                a, b = MAT_MODE[i]
                v = VEC_MODE[i]
                size2 = (self.K, res[v]) if self.dynamic else (res[v], 1)
                if planes[i].shape[1] == 0:
                    # empty groups are re-created at the new size (the dynamic net writes zeros, :402-408; an interpolate of a
                    # 0-channel tensor gives the same empty tensor for the static one)
                    planes[i] = nn.Parameter(planes[i].data.new_zeros(1, comps[i], res[b], res[a]))
                    seconds[i] = nn.Parameter(seconds[i].data.new_zeros(1, comps[i], *size2))
                    continue
                planes[i] = nn.Parameter(F.interpolate(planes[i].data, size=(res[b], res[a]), mode="bilinear", align_corners=True))
                seconds[i] = nn.Parameter(F.interpolate(seconds[i].data, size=size2, mode="bilinear", align_corners=True))
        self.update_stepSize(res)
        self.gridSize = self.gridSize.to(dp[0].device)
        self.struct_version = getattr(self, "struct_version", 0) + 1

    # ---- regulariser terms of nlf/regularizers/tensorf.py:35-96 (tensorf_base.py:1024-1057, tensorf_dynamic.py:246-286)
    def density_L1(self):
        dp, d2, _, _ = self.tables()
        total = 0
        for i in range(3):
            if dp[i].shape[1] == 0:
                continue
            total = total + torch.mean(torch.abs(dp[i])) + torch.mean(torch.abs(d2[i]))
        return total

    def TV_loss_density(self, reg):
        dp, _, _, _ = self.tables()
        total = 0
        for i in range(3):
            if dp[i].shape[1] == 0:
                continue
            total = total + reg(dp[i]) * 1e-2
        return total

    def TV_loss_app(self, reg):
        dp, _, ap, _ = self.tables()
        total = 0
        for i in range(3):
            if (dp[i].shape[1] if self.dynamic else ap[i].shape[1]) == 0:
                continue
            total = total + reg(ap[i]) * 1e-2
        return total

    def tables(self):
        if self.dynamic:
            return self.density_plane_space, self.density_plane_time, self.app_plane_space, self.app_plane_time
        return self.density_plane, self.density_line, self.app_plane, self.app_line

    def set_iter(self, i):
        """TensorBase.set_iter (tensorf_base.py:509-553), the up-sampling half: in training mode, at the iterations of
        `upsamp_list`, re-sample every table to the next grid of the schedule and ask for an optimiser reset.  The alpha-mask
        update / aabb shrink of `update_AlphaMask_list` (:517-529) is not mirrored."""
        self.cur_iter = i
        if not self.training:
            return
        self.needs_opt_reset = False
        if i in self.upsamp_list and len(self.N_voxel_list) > 0:
            if self.use_grid_size_upsample:
                if len(self.N_voxel_list[0]) == 0:
                    return
                reso = [self.N_voxel_list[k].pop(0) for k in range(3)]
            else:
                reso = n_to_reso(self.N_voxel_list.pop(0), self.aabb.detach().cpu())
            self.upsample_volume_grid(reso)
            if self.lr_upsample_reset:
                self.needs_opt_reset = True


class _Color(nn.Module):
    def __init__(self, sig, grid):
        super().__init__()
        self.net = _Tensorf(sig, grid)

    def set_iter(self, i):
        self.net.set_iter(i)


def default_grid(sig: Signature) -> List[int]:
    """Initial grid of the reference constructor: N_to_reso(N_voxel_init, aabb) (tensorf_base.py:156-159)."""
    net = sig.model_cfg.color.net
    if "grid_size" in net:
        return list(net.grid_size.start)
    return n_to_reso(int(net.N_voxel_init), torch.tensor(net.aabb))


def scale_density(sd: Dict[str, torch.Tensor], gain: float) -> Dict[str, torch.Tensor]:
    """"Trained-like" variant: scale the sigma space planes so transmittance
    saturates along a ray instead of leaving every sample nearly transparent."""
    out = {}
    for k, v in sd.items():
        if ("density_plane" in k) and ("time" not in k):
            out[k] = v * gain
        else:
            out[k] = v
    return out


def scale_appearance(sd: Dict[str, torch.Tensor], gain: float) -> Dict[str, torch.Tensor]:
    """Scale both factors of every appearance table (planes and lines / time planes): the reference initialises them at
    0.1 * N(0,1), so their products barely move the shaded colour; trained scenes have O(1) features."""
    return {k: (v * gain if ".app_" in k else v) for k, v in sd.items()}


def seeded_state_dict(sig: Signature, grid: Optional[Sequence[int]] = None, seed: int = 0, density_gain: float = 1.0,
                      prefix: str = "model.", app_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Reference-style random initialisation of every parameter, deterministic in ``seed`` (CPU generator),
    keyed like ``RenderLightfield.state_dict()`` (``model.embedding_model...``, ``model.color_model.net...``)."""
    grid = list(grid) if grid is not None else default_grid(sig)
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        emb = _Embedding(sig.mlp_layer_shapes, sig.color_views, sig.color_embedding_index, sig.net_index, sig.pre_layer_shapes)
        if sig.color_views > 0:  # the reference initialises the table with zeros (identity transform): give the tests something to see
            emb.embeddings[sig.color_embedding_index].color_embedding.data.normal_(0.0, 0.5)
        col = _Color(sig, grid)
    sd = {}
    for k, v in emb.state_dict().items():
        sd[f"{prefix}embedding_model.{k}"] = v.detach().clone()
    for k, v in col.state_dict().items():
        sd[f"{prefix}color_model.{k}"] = v.detach().clone()
    if density_gain != 1.0:
        sd = scale_density(sd, density_gain)
    if app_gain != 1.0:
        sd = scale_appearance(sd, app_gain)
    return sd
