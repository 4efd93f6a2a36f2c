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


class _AlphaGrid:
    """Binary occupancy volume over an aabb (the role of AlphaGridMask, utils/tensorf_utils.py:459-484): `alpha` is [Z,Y,X]."""

    def __init__(self, aabb, alpha):
        self.aabb = aabb.clone()
        self.volume = alpha.view(1, 1, *alpha.shape[-3:])
        self.gridSize = torch.tensor([alpha.shape[-1], alpha.shape[-2], alpha.shape[-3]], dtype=torch.long)

    def sample_alpha(self, pts):
        import torch.nn.functional as F

        u = (pts - self.aabb[0]) * (1.0 / (self.aabb[1] - self.aabb[0]) * 2) - 1
        return F.grid_sample(self.volume, u.view(1, -1, 1, 1, 3), align_corners=True).view(-1)


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
        self.update_AlphaMask_list = [int(v) for v in (net.get("update_AlphaMask_list", []) or [])]
        self.alphaMask_thres = float(net.get("alpha_mask_thre", 0.001))
        self.fea2denseAct = net.get("fea2denseAct", "softplus")
        self.density_shift = float(net.get("density_shift", -10.0))
        self.total_num_frames = int(c.num_frames) if c.num_frames > 0 else 1
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

    # ---- occupancy pruning: alpha-mask update and aabb shrink (tensorf_base.py:379-429,1190-1232; dynamic net
    # tensorf_dynamic.py:443-541,618-643).  Host-side schedule steps that run once or twice per training, so they are torch ops
    # on the reference-layout tables; the render path only ever sees their result (a smaller aabb, cropped tables): the
    # reference disables the mask lookup in its forward (`if self.alphaMask is not None and False`, tensorf_dynamic.py:707).
    def _density_feature(self, u, tau=None):
        """Sum over groups and channels of space-plane x second-factor samples at normalised coordinates u [M,3] (tau [M]:
        normalised keyframe time of the dynamic net)."""
        import torch.nn.functional as F

        dp, d2, _, _ = self.tables()
        M = u.shape[0]
        feat = torch.zeros((M,), device=u.device, dtype=u.dtype)
        for i in range(3):
            if dp[i].shape[1] == 0:
                continue
            a, b = MAT_MODE[i]
            v = VEC_MODE[i]
            plane = F.grid_sample(dp[i], torch.stack((u[:, a], u[:, b]), -1).view(1, M, 1, 2), align_corners=True).view(-1, M)
            if self.dynamic:
                coord = torch.stack((u[:, v], tau), -1)
            else:
                coord = torch.stack((torch.zeros_like(u[:, v]), u[:, v]), -1)
            second = F.grid_sample(d2[i], coord.view(1, M, 1, 2), align_corners=True).view(-1, M)
            feat = feat + torch.sum(plane * second, dim=0)
        return feat

    def _feature2density(self, feat):
        import torch.nn.functional as F

        if self.fea2denseAct == "softplus":
            return F.softplus(feat + self.density_shift)
        if self.fea2denseAct == "relu":
            return F.relu(feat)
        return F.relu(torch.abs(feat))

    def compute_alpha(self, pts, base_times=None, length: float = 0.01):
        """1 - exp(-sigma * length) at world points pts [M,3] (tensorf_base.py:489-507; dynamic: tensorf_dynamic.py:618-643 at
        keyframe times base_times [M]).  The static net skips points an earlier mask marks empty."""
        aabb = self.aabb.to(pts.device)
        inv = 2.0 / (aabb[1] - aabb[0])
        sigma = torch.zeros(pts.shape[0], device=pts.device, dtype=pts.dtype)
        keep = torch.ones(pts.shape[0], dtype=torch.bool, device=pts.device)
        if not self.dynamic and self.alphaMask is not None:
            keep = self.alphaMask.sample_alpha(pts) > 0
        if keep.any():
            u = (pts[keep] - aabb[0]) * inv - 1
            tau = None
            if self.dynamic:
                Fr = self.total_num_frames
                tau = (base_times[keep] * ((Fr - 1) / Fr) + 0.5 / self.K) * 2 - 1  # normalize_time_coord (:615-616)
            sigma[keep] = self._feature2density(self._density_feature(u, tau))
        return 1 - torch.exp(-sigma * length)

    @torch.no_grad()
    def getDenseAlpha(self, gridSize):
        gx, gy, gz = [int(v) for v in gridSize]
        dev = self.tables()[0][0].device
        aabb = self.aabb.to(dev)
        lin = [torch.linspace(0, 1, n) for n in (gx, gy, gz)]
        samples = torch.stack(torch.meshgrid(*lin, indexing="ij"), -1).to(dev)
        dense_xyz = aabb[0] * (1 - samples) + aabb[1] * samples
        flat = dense_xyz.view(-1, 3)
        if not self.dynamic:
            return self.compute_alpha(flat).view(gx, gy, gz), dense_xyz
        # the dynamic net takes the maximum over all frames; a frame's keyframe time follows this function's own snap
        # (tensorf_dynamic.py:516-521: scale (F-1)/F, not the K(F-1)/F of get_base_time)
        alpha = torch.zeros(gx, gy, gz, device=dev)
        Fr = self.total_num_frames
        tsf = (Fr - 1) / Fr
        import numpy as np
        for t in np.linspace(0, 1, Fr):
            times = torch.ones(flat.shape[0], device=dev) * t
            base = torch.round((times * tsf).clamp(0.0, self.K - 1)) * (1.0 / tsf)
            alpha = torch.maximum(alpha, self.compute_alpha(flat, base).view(gx, gy, gz))
        return alpha, dense_xyz

    @torch.no_grad()
    def updateAlphaMask(self, gridSize=(200, 200, 200)):
        """Dense occupancy -> 3x3x3 dilation -> threshold -> mask volume (kept as `alphaMask`) and the bounding box of the
        occupied voxels (returned)."""
        import torch.nn.functional as F

        g = [int(v) for v in gridSize]
        alpha, dense_xyz = self.getDenseAlpha(g)
        dense_xyz = dense_xyz.transpose(0, 2).contiguous()
        alpha = alpha.clamp(0, 1).transpose(0, 2).contiguous()[None, None]
        alpha = F.max_pool3d(alpha, kernel_size=3, padding=1, stride=1).view(g[::-1])
        alpha = (alpha >= self.alphaMask_thres).to(alpha.dtype)
        self.alphaMask = _AlphaGrid(self.aabb.to(alpha.device), alpha)
        occupied = dense_xyz[alpha > 0.5]
        return torch.stack((occupied.amin(0), occupied.amax(0)))

    @torch.no_grad()
    def shrink(self, new_aabb):
        """Crop every table to the texel range covering new_aabb and move the aabb to the cropped grid's corners."""
        dp, d2, ap, a2 = self.tables()
        dev = dp[0].device
        aabb, grid = self.aabb.to(dev), self.gridSize.to(dev)
        units = (aabb[1] - aabb[0]) / (grid - 1)
        lo = torch.round(torch.round((new_aabb[0].to(dev) - aabb[0]) / units)).long()
        hi = torch.minimum(torch.round((new_aabb[1].to(dev) - aabb[0]) / units).long() + 1, grid)
        for i in range(3):
            a, b = MAT_MODE[i]
            v = VEC_MODE[i]
            for planes, seconds in ((dp, d2), (ap, a2)):
                planes[i] = nn.Parameter(planes[i].data[..., lo[b]:hi[b], lo[a]:hi[a]])
                seconds[i] = nn.Parameter(seconds[i].data[..., :, lo[v]:hi[v]] if self.dynamic else seconds[i].data[..., lo[v]:hi[v], :])
        if self.alphaMask is None or not torch.all(self.alphaMask.gridSize.to(dev) == grid):
            lo_r, hi_r = lo / (grid - 1), (hi - 1) / (grid - 1)
            new_aabb = torch.stack(((1 - lo_r) * aabb[0] + lo_r * aabb[1], (1 - hi_r) * aabb[0] + hi_r * aabb[1]))
        self.aabb = new_aabb.to(self.aabb.device).to(self.aabb.dtype)
        self.update_stepSize((hi - lo).tolist())
        self.gridSize = self.gridSize.to(dev)
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
        `update_AlphaMask_list`, rebuild the occupancy mask (and shrink the aabb / crop the tables the first time); at those of
        `upsamp_list`, re-sample every table to the next grid of the schedule and ask for an optimiser reset."""
        self.cur_iter = i
        if not self.training:
            return
        self.needs_opt_reset = False
        if i in self.update_AlphaMask_list:  # pruning (:517-529): mask at the grid's resolution, capped at 200^3
            reso = tuple(int(v) for v in self.gridSize.tolist())
            if reso[0] > 200:
                reso = (200, 200, 200)
            new_aabb = self.updateAlphaMask(reso)
            if i == self.update_AlphaMask_list[0]:
                self.shrink(new_aabb)
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
