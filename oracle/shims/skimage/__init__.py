"""Import shim: occdepth/data/utils/fusion.py:15 does `from skimage import measure` for its mesh export only
(marching cubes); nothing on the projection path (vox2pix) touches it.  Test infrastructure, like the other shims."""
from . import measure  # noqa: F401
