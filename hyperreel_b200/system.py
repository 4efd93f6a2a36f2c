"""Render-path surface of the reference's ``INRSystem`` (nlf/__init__.py:278-502).

Only what sits on the hot path is kept: construction from the full config (``cfg.model``, ``cfg.training``,
``cfg.dataset``), ``forward`` / ``render`` / ``run_chunked`` with the reference's chunk selection,
``load_state_dict`` with the reference's grid-size fix-up (nlf/__init__.py:433-479), including Lightning
checkpoints whose keys carry the ``render_fn.`` prefix, and -- SURVEY.md section 8 row f1 -- ``configure_optimizers`` /
``training_step`` (nlf/__init__.py:504-523, 634-709) over the differentiable path: image loss, manual optimisation with
one Adam per optimiser group, the TensoRF regulariser (L1 + TV on the tables, nlf/regularizers/tensorf.py:35-96) and the grid
up-sampling / occupancy-pruning schedule with its optimiser reset (tensorf_base.py:379-429,509-553,1151-1232).  Visualisers and
datasets are out of scope.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from .config import Cfg, epochs_to_iters, to_cfg
from .models import model_dict
from .rendering import render_chunked, render_fn_dict


class TVLoss(nn.Module):
    """nlf/regularizers/tensorf.py:14-32: 2 * (mean squared row difference + mean squared column difference) / batch."""

    def forward(self, x):
        count_h = x[:, :, 1:, :].size(1) * x[:, :, 1:, :].size(2) * x[:, :, 1:, :].size(3)
        count_w = x[:, :, :, 1:].size(1) * x[:, :, :, 1:].size(2) * x[:, :, :, 1:].size(3)
        h_tv = torch.pow(x[:, :, 1:, :] - x[:, :, :-1, :], 2).sum()
        w_tv = torch.pow(x[:, :, :, 1:] - x[:, :, :, :-1], 2).sum()
        return 2 * (h_tv / count_h + w_tv / count_w) / x.size(0)


class TensoRFRegularizer:
    """The TensoRF regulariser of the reference (nlf/regularizers/tensorf.py:35-96; conf/experiment/regularizers/tensorf/
    *.yaml): L1 on the density tables plus total variation on the space planes, the TV weights decaying by `lr_factor` per
    use exactly like the reference (which scales the *running* weights but multiplies by the configured ones, :79-88)."""

    def __init__(self, cfg):
        import math

        self.cfg = to_cfg(cfg)
        self.tvreg = TVLoss()
        self.cur_iter = 0
        self.update_AlphaMask_list = list(self.cfg.get("update_AlphaMask_list", []))
        self.lr_factor = float(self.cfg.lr_decay_target_ratio) ** (1.0 / float(self.cfg.n_iters))
        self.total_num_tv_iters = (int(self.cfg.total_num_tv_iters) if "total_num_tv_iters" in self.cfg else
                                   int(round(math.log(1e-4) / math.log(float(self.cfg.lr_decay_target_ratio)) * float(self.cfg.n_iters))))
        self.L1_reg_weight = float(self.cfg.L1_weight_initial)
        self.TV_weight_density = float(self.cfg.TV_weight_density)
        self.TV_weight_app = float(self.cfg.TV_weight_app)

    def loss(self, tensorf):
        total = 0.0
        if self.L1_reg_weight > 0:
            total = total + self.L1_reg_weight * tensorf.density_L1()
        if self.cur_iter > self.total_num_tv_iters:
            return total
        loss_tv = 0.0
        if self.TV_weight_density > 0:
            self.TV_weight_density *= self.lr_factor
            loss_tv = tensorf.TV_loss_density(self.tvreg) * float(self.cfg.TV_weight_density)
            total = total + loss_tv
        if self.TV_weight_app > 0:
            self.TV_weight_app *= self.lr_factor
            loss_tv = loss_tv + tensorf.TV_loss_app(self.tvreg) * float(self.cfg.TV_weight_app)  # the density term is counted twice, as in :85-88
            total = total + loss_tv
        return total

    def set_iter(self, iteration):
        self.cur_iter = iteration
        if len(self.update_AlphaMask_list) > 0 and self.cur_iter == self.update_AlphaMask_list[0]:
            self.L1_reg_weight = float(self.cfg.L1_weight_rest)


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
        # regularisers (nlf/__init__.py:396-407): only the TensoRF one touches this path's parameters
        self.regularizers = []
        for key, rcfg in (self.cfg.get("regularizers", Cfg()) or Cfg()).items():
            if rcfg.get("type") == "tensorf":
                self.regularizers.append(TensoRFRegularizer(rcfg))
            else:
                raise NotImplementedError(f"regularizer '{rcfg.get('type')}' is outside the fused path's scope")
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

    # ---- nlf/__init__.py:504-523 + utils/__init__.py:49-76 (get_optimizer): one Adam(betas=(0.9, 0.99), eps=1e-8) per group
    OPT_DEFAULTS = {"color": 0.02, "color_impl": 0.001, "embedding_impl": 0.00075}  # conf/experiment/training/*_tensorf.yaml

    def optimizer_groups(self):
        """Parameters by the reference's `opt_group` names: the VM tables ('color'), basis_mat ('color_impl'), the sample
        net ('embedding_impl') (nlf/nets/tensorf_base.py opt_group dict, nlf/embedding/ray.py net group)."""
        model = self.render_fn.model
        net = model.color_model.net
        tables = [p for n, p in net.named_parameters() if "plane" in n or "line" in n]
        impl = [p for n, p in net.named_parameters() if "basis_mat" in n]
        return {"color": tables, "color_impl": impl, "embedding_impl": list(model.embedding_model.parameters())}

    def configure_optimizers(self):
        training = self.cfg.get("training", Cfg())
        ocfg = training.get("optimizers", Cfg())
        self._optimizers = []
        self._opt_struct = self.render_fn.model.color_model.net.struct_version
        for key, params in self.optimizer_groups().items():
            params = [p for p in params if p.numel() > 0]
            if not params:
                continue
            oc = ocfg.get(key, Cfg())
            if oc.get("optimizer", "adam") != "adam":
                raise NotImplementedError("only the reference's default optimizer (adam) is mirrored")
            self._optimizers.append(torch.optim.Adam(params, lr=float(oc.get("lr", self.OPT_DEFAULTS[key])), eps=1e-8,
                                                     weight_decay=float(oc.get("weight_decay", 0)), betas=(0.9, 0.99)))
        return self._optimizers

    def set_train_iter(self, train_iter: int):
        """The per-iteration hook of the reference's loop (nlf/__init__.py:592-632): the colour net's schedule (grid
        up-sampling, tensorf_base.py:509-553) and the regularisers' (`L1_weight_rest`), then an optimiser rebuild when the
        tables were re-created (`lr_upsample_reset`, nlf/__init__.py:541-578 -- here every group restarts, Adam state of the old
        Parameter objects is meaningless for the new ones)."""
        model = self.render_fn.model
        model.color_model.set_iter(int(train_iter))
        for reg in self.regularizers:
            reg.set_iter(int(train_iter))
        net = model.color_model.net
        if getattr(net, "needs_opt_reset", False) or (getattr(self, "_opt_struct", None) not in (None, net.struct_version)):
            self.configure_optimizers()
            net.needs_opt_reset = False

    def training_step(self, batch, batch_idx: int = 0, train_iter: Optional[int] = None):
        """One iteration of nlf/__init__.py:634-709 on this path: batch {'coords' [N,C], 'rgb' [N,3], optional 'weight'};
        loss = MSE(rgb_pred * w, rgb * w) (losses.py 'mse') + regularisers, manual optimisation (zero_grad / backward / step
        per group).  `train_iter` (optional) drives the schedules like the reference's `set_train_iter`."""
        self.train()
        if train_iter is not None:
            self.set_train_iter(train_iter)
        if not getattr(self, "_optimizers", None):
            self.configure_optimizers()
        coords, rgb = batch["coords"], batch["rgb"]
        weight = batch.get("weight", None)
        results = self(coords)
        pred = results["rgb"]
        if weight is not None:
            loss = torch.mean((pred * weight - rgb * weight) ** 2)
        else:
            loss = torch.mean((pred - rgb) ** 2)
        for reg in self.regularizers:  # nlf/__init__.py:677-683 (loss weight 1: `exponential_decay` with decay 1.0)
            loss = loss + reg.loss(self.render_fn.model.color_model.net)
        for opt in self._optimizers:
            opt.zero_grad(set_to_none=True)
        loss.backward()
        for opt in self._optimizers:
            opt.step()
        with torch.no_grad():
            psnr = -10.0 * torch.log10(torch.mean((pred.detach() - rgb) ** 2))  # metrics.py:37-45
        return {"train/loss": loss.detach(), "train/psnr": psnr}

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
