"""Drop-in alias: `import starster` resolves to the B200-native implementation (starst3r_b200)."""
from starst3r_b200 import *  # noqa: F401,F403
from starst3r_b200 import Scene, gs, match, __version__  # noqa: F401


def __getattr__(name):
    import starst3r_b200
    return getattr(starst3r_b200, name)
