"""Drop-in alias: `import starster` resolves to the B200-native implementation (starst3r_b200).

Mirrors starster/__init__.py:1-9 of the reference: `Mast3rModel`, `gs`, and the star-exports of `image`,
`reconstruct`, `scene`, `utils`; the reference's submodules (`starster.gs`, `starster.reconstruct`, `starster.scene`,
`starster.image`, `starster.utils`) resolve to the corresponding starst3r_b200 modules, so
`from starster.reconstruct import reconstruct_scene` keeps working."""
import importlib
import sys

import starst3r_b200
from starst3r_b200 import *  # noqa: F401,F403
from starst3r_b200 import Scene, gs, match, __version__  # noqa: F401

for _name in ("gs", "image", "match", "reconstruct", "scene", "utils"):
    sys.modules[__name__ + "." + _name] = importlib.import_module("starst3r_b200." + _name)
del _name


def __getattr__(name):
    return getattr(starst3r_b200, name)
