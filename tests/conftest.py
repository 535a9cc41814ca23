import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATA = os.path.join(ROOT, "data", "haarcascades")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def cascade_path(name: str) -> str:
    return os.path.join(DATA, f"haarcascade_{name}.xml")


# all 19 cascade files the reference ships; the first six are BASELINE.json's
ALL_CASCADES = ["frontalface_alt", "frontalface_default", "frontalface_alt_tree", "eye", "profileface",
                "fullbody", "frontalface_alt2", "eye_tree_eyeglasses", "mcs_nose",
                "lefteye_2splits", "righteye_2splits", "lowerbody", "upperbody", "mcs_eyepair_big",
                "mcs_eyepair_small", "mcs_lefteye", "mcs_mouth", "mcs_righteye", "mcs_upperbody"]
# the nine round-1 cascades (regression fixtures tests/golden/refsi_*.npz exist for these)
CORE_CASCADES = ALL_CASCADES[:9]


@pytest.fixture(scope="session")
def gpu_ctx():
    import clfacedetection_b200 as clfd
    ctx = clfd.Context(0)   # raises (no fallback) when there is no GPU / no built library
    yield ctx
    ctx.close()


_oracle_cache = {}


def oracle_cascade(name):
    import oracle
    if name not in _oracle_cache:
        _oracle_cache[name] = oracle.Cascade(cascade_path(name))
    return _oracle_cache[name]
