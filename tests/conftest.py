import importlib
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle_lib  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def _cuda_available() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def fm():
    """the product package (directory name has a hyphen)"""
    return importlib.import_module("bachelor-thesis_b200")


@pytest.fixture(scope="session")
def oracle():
    if not os.path.exists(oracle_lib.ORACLE_SO):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def ref():
    if not oracle_lib.ref_available():
        pytest.skip("oracle/_ref/libfluidref.so not built (needs /root/reference at build time)")
    return oracle_lib.Ref()


@pytest.fixture(scope="session")
def gpu_ctx_factory(fm):
    """GPU tests must run on the CUDA library; a missing device or .so is an error, not a skip."""
    if not _cuda_available():
        pytest.fail("-m gpu tests need a CUDA device")
    made = []

    def make(w, h):
        c = fm.Context(w, h)
        made.append(c)
        return c

    yield make
    for c in made:
        c.close()


def golden_camera(name="camera_default_16x9"):
    import json
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        d = json.load(f)
    return {k: np.array(v, dtype=np.float32) for k, v in d["float32"].items()}


@pytest.fixture(scope="session")
def default_camera():
    """matrices dumped from the reference's own Camera3D/CameraController3D (tests/golden/make_golden.py)"""
    return golden_camera()
