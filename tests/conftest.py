import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # keep the in-tree CUDA library and the oracle in step with their sources (no-ops when up to date)
    import subprocess
    for d in (os.path.join(ROOT, "syzgydb_b200", "csrc"), os.path.join(ROOT, "oracle")):
        try:
            subprocess.run(["make", "-C", d, "-j", str(os.cpu_count() or 2)], check=False, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, timeout=900)
        except Exception:
            pass


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
