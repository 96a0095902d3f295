import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _build_native():
    """Make sure the native pieces exist and are newer than their sources BEFORE any test module imports the package
    (importing it loads librtdd.so and fails loudly without it).  The recipe is loaded by path for the same reason."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("rtdd_build_recipe", os.path.join(ROOT, "realtimedepthdiffusion_b200", "build.py"))
    recipe = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(recipe)
    recipe.build_all()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    _build_native()


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
