"""The documented drop-in order: `occdepth_b200.install_as_occdepth()` BEFORE the
reference's imports, after which the body of the reference scripts must import -- `occdepth.models*` from this
package, everything else (`occdepth.data.*`, `occdepth.loss.*`) still from the reference tree.  Runs in a subprocess
so that the test session's own sys.modules stay untouched."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WITH_REFERENCE = r"""
import sys
sys.path[:0] = [%(root)r, %(shims)r, %(ref)r]
import occdepth_b200
occdepth_b200.install_as_occdepth()
# the import block of occdepth/scripts/generate_output.py:1-5 and eval.py (minus the third-party trainer)
from occdepth.models.OccDepth import OccDepth
import occdepth.loss.sscMetrics as sscMetrics
import occdepth.data.utils.helpers as helpers
import occdepth.data.semantic_kitti.collate as collate
from occdepth.models.unet3d_kitti import UNet3D
from occdepth.models.SFA import SFA
import occdepth
assert OccDepth.__module__ == "occdepth_b200.models.OccDepth", OccDepth.__module__
assert UNet3D.__module__ == "occdepth_b200.models.unet3d_kitti" and SFA.__module__ == "occdepth_b200.models.SFA"
assert "reference" in sscMetrics.__file__ or %(ref)r in sscMetrics.__file__, sscMetrics.__file__
assert hasattr(helpers, "vox2pix") and hasattr(collate, "collate_fn")
assert hasattr(occdepth, "__path__")
print("OK")
"""

WITHOUT_REFERENCE = r"""
import sys
sys.path.insert(0, %(root)r)
import occdepth_b200
occdepth_b200.install_as_occdepth()
from occdepth.models.OccDepth import OccDepth
from occdepth.models.flosp_depth.flosp_depth import FlospDepth
assert OccDepth.__module__ == "occdepth_b200.models.OccDepth"
try:
    import occdepth.data.utils.helpers
except ModuleNotFoundError as e:
    assert "is not a package" not in str(e), e
else:
    raise SystemExit("occdepth.data must not resolve without the reference tree")
print("OK")
"""


def _run(code):
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd="/tmp")


@pytest.mark.reference
def test_install_then_reference_imports():
    from oracle import ref_import
    r = _run(WITH_REFERENCE % {"root": ROOT, "shims": os.path.join(ROOT, "oracle", "shims"), "ref": ref_import.REF_ROOT})
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


def test_install_without_reference_tree():
    r = _run(WITHOUT_REFERENCE % {"root": ROOT})
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


def test_frame_pipeline_refuses_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from occdepth_b200.serving import FramePipeline
    with pytest.raises(RuntimeError, match="CUDA"):
        FramePipeline(None, (1, 2, 3, 8, 8), (2, 8, 1, 2), (2, 8, 1))
