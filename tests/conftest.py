import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    # The oracle runs on the host CPU.  oneDNN convolutions stop scaling beyond ~16 threads and can become pathologically slow
    # on the 128-thread GPU boxes (one 256^2 U-Net evaluation: 0.22 s at 16 threads, 61 s at 128 - DESIGN.md §5), so cap them.
    try:
        import torch
        torch.set_num_threads(min(16, os.cpu_count() or 1))
    except Exception:
        pass
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_shim
    if ref_shim.available():
        return
    skip = pytest.mark.skip(reason="reference tree not present (GPU box)")
    for it in items:
        if "reference" in it.keywords:
            it.add_marker(skip)
