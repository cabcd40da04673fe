"""starst3r_b200 — B200-native (sm_100a) hot paths of phuang1024/Starst3r behind its Python API."""
__version__ = "0.1.0"

from . import match  # noqa: F401
