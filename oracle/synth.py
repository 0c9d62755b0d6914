"""Synthetic-input helpers live in /synthetic.py (shared by bench.py and the tests); re-exported here so that oracle
users keep one import point."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from synthetic import *  # noqa: E402,F401,F403
from synthetic import Cfg, feature_hw, kitti_calib, kitti_indices, occdepth_cfg, random_indices  # noqa: E402,F401
from synthetic import randomize_bn_, seed_weights_, vox2pix  # noqa: E402,F401
