"""starst3r_b200 — B200-native (sm_100a) hot paths of phuang1024/Starst3r behind its Python API
(`Scene`, `reconstruct_scene`, `gs.*`).  See DESIGN.md."""
__version__ = "0.1.0"

from . import gs, match  # noqa: F401
from .image import load_image, load_images, prepare_images_for_mast3r, process_image  # noqa: F401
from .scene import Scene  # noqa: F401
from .utils import interp_se3, interp_se3_path  # noqa: F401


def __getattr__(name):
    if name in ("reconstruct_scene", "run_sparse_ga", "sparse_scene_optimizer_slam"):
        from . import reconstruct
        return getattr(reconstruct, name)
    raise AttributeError(name)
