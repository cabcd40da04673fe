"""starst3r_b200 — B200-native (sm_100a) hot paths of phuang1024/Starst3r behind its Python API
(`Scene`, `reconstruct_scene`, `gs.*`).  See DESIGN.md."""
__version__ = "0.1.0"

from . import gs, match  # noqa: F401
from .image import load_image, load_images, make_pair_indices, prepare_images_for_mast3r, process_image  # noqa: F401
from .scene import Scene  # noqa: F401
from .utils import interp_se3, interp_se3_path  # noqa: F401


class _Mast3rModelUnavailable:
    """Stand-in for `starster.Mast3rModel` (starster/__init__.py:3: `mast3r.model.AsymmetricMASt3R`) when the MASt3R
    network package is not importable.  The network is outside this package's scope: put the
    reference's `mast3r`, `mast3r/dust3r` and `mast3r/dust3r/croco` directories on sys.path (as main.py:6-8 does) and
    the real class is returned instead."""

    @classmethod
    def from_pretrained(cls, *args, **kwargs):
        raise ImportError("starster.Mast3rModel: the `mast3r` package (mast3r.model.AsymmetricMASt3R) is not importable; "
                          "add the reference's mast3r, mast3r/dust3r and mast3r/dust3r/croco directories to sys.path")

    def __init__(self, *args, **kwargs):
        self.from_pretrained()


def __getattr__(name):
    if name in ("reconstruct_scene", "run_sparse_ga", "sparse_scene_optimizer_slam"):
        from . import reconstruct
        return getattr(reconstruct, name)
    if name == "reconstruct":
        import importlib
        return importlib.import_module(".reconstruct", __name__)
    if name == "Mast3rModel":
        try:
            from mast3r.model import AsymmetricMASt3R
            return AsymmetricMASt3R
        except Exception:     # not installed, or one of its own dependencies is missing
            return _Mast3rModelUnavailable
    raise AttributeError(name)
