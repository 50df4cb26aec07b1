import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_cuda():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """Drop the engines (captured CUDA graphs, workspaces) while the CUDA context is still healthy instead
    of leaving them to interpreter finalisation."""
    import gc
    try:
        import torch
        if not (torch.cuda.is_available() and torch.cuda.is_initialized()):
            return
        from video_description_with_spatial_temporal_attention_b200 import engine
        engine._release_graphs()
        gc.collect()
        torch.cuda.synchronize()
    except Exception:
        pass
