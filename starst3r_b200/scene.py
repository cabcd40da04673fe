"""`Scene` with the attribute / method surface of starster/scene.py:18-183."""
import tempfile
from typing import Optional

import torch

from . import gs as _gs

__all__ = ("Scene",)


class Scene:
    """Holds the MASt3R reconstruction (cameras, dense points) and the 3DGS state of one scene.
    Attributes mirror starster/scene.py:47-77; `gaussians`, `optimizers`, `strategy`, `strategy_state` appear after
    init_3dgs() exactly as in the reference (gs.py:20-45)."""

    def __init__(self, cache_dir: Optional[str] = None, device="cuda"):
        self.device = device
        self._own_cache = cache_dir is None      # a private cache: its device-resident memo dies with the scene
        self.cache_dir = cache_dir if cache_dir is not None else tempfile.mkdtemp()
        self.raw_imgs = []
        self.imgs = []
        self.dense_pts = []
        self.dense_cols = []
        self.c2w = None
        self.intrinsics = None
        self.optim_params = None
        self.gs_params = None
        self.gs_optims = None
        self.gs_strategy = None
        self.gs_state = None

    def __del__(self):
        if getattr(self, "_own_cache", False):
            try:
                from .reconstruct import clear_cache
                clear_cache(self.cache_dir)
            except Exception:       # noqa: BLE001 - interpreter shutdown
                pass

    @property
    def dense_pts_flat(self):
        assert self.dense_pts, "No dense points available."
        return torch.cat(self.dense_pts, dim=0)

    @property
    def dense_cols_flat(self):
        assert self.dense_cols, "No dense colors available."
        return torch.cat(self.dense_cols, dim=0)

    @property
    def w2c(self) -> torch.Tensor:
        assert self.c2w is not None, "No c2w matrix available."
        return torch.inverse(self.c2w)

    def add_images(self, model, imgs, conf_thres=1.5):
        """scene.py:97-155: re-runs the reconstruction over all images and refreshes cameras + dense points."""
        from .reconstruct import reconstruct_scene
        self.raw_imgs.extend(imgs)
        filelist = [f"{i}.png" for i in range(len(self.raw_imgs))]
        scene, optim_params = reconstruct_scene(model, self.raw_imgs, filelist, self.device,
                                                optim_params=self.optim_params, tmpdir=self.cache_dir)
        self.optim_params = optim_params
        self.imgs.extend(scene.imgs[len(self.imgs):])
        self.c2w = scene.cam2w
        self.intrinsics = scene.intrinsics
        pts, _, confs = scene.get_dense_pts3d(clean_depth=True)
        self.dense_pts, self.dense_cols = [], []
        for i in range(len(scene.imgs)):
            mask = (confs[i] > conf_thres).reshape(-1).cpu()
            colors = torch.tensor(scene.imgs[i]).reshape(-1, 3)
            self.dense_pts.append(pts[i][mask])
            self.dense_cols.append(colors[mask])

    def init_3dgs(self, init_scale=3e-3, lr=1e-3):
        _gs.init_3dgs(self, init_scale, lr)

    def render_3dgs(self, w2c, intrinsics, width, height):
        return _gs.render_3dgs(self, w2c, intrinsics, width, height)

    def render_3dgs_original(self, width, height):
        return _gs.render_3dgs_original(self, width, height)

    def run_3dgs_optim(self, iters: int, enable_pruning: bool = False, loss_ssim_fac=0.2, loss_opacity_fac=0.01,
                       loss_scale_fac=0.01, verbose: bool = False):
        return _gs.run_3dgs_optim(self, iters, enable_pruning, loss_ssim_fac, loss_opacity_fac, loss_scale_fac, verbose)
