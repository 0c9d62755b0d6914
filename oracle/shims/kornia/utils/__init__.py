from . import grid  # noqa: F401
from .grid import create_meshgrid3d  # noqa: F401
