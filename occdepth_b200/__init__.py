"""occdepth_b200 -- B200 (sm_100a) native forward hot path of OccDepth.

Module surface mirrors the reference's `occdepth.models` (same class names, constructor signatures, forward
inputs/outputs and state_dict keys); `install_as_occdepth()` registers this package under the reference's
import names so its train/eval scripts import it unchanged.
"""
__version__ = "0.1.0"


def install_as_occdepth():
    """Expose occdepth_b200.models.* as occdepth.models.* (the reference's import paths)."""
    import importlib
    import sys
    import types

    names = ["SFA", "DDR", "modules", "CRP3D", "unet3d_kitti", "unet3d_nyu", "unet2d", "OccDepth"]
    root = sys.modules.setdefault("occdepth", types.ModuleType("occdepth"))
    models = importlib.import_module("occdepth_b200.models")
    sys.modules["occdepth.models"] = models
    root.models = models
    for n in names:
        try:
            m = importlib.import_module("occdepth_b200.models." + n)
        except ImportError:
            continue
        sys.modules["occdepth.models." + n] = m
    return models
