import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_import
    have_ref = ref_import.available()
    skip_ref = pytest.mark.skip(reason="reference tree (/root/reference) not present on this machine")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
