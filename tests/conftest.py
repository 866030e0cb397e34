import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
CHAINS = ["c6", "c7", "c6_perturbed", "c7_perturbed", "random_a", "random_b", "random_c", "random_d"]

# fp64 tolerance of BASELINE.json's north_star: <= 1e-10 relative (kinematics, torque, regressor, inertia),
# per output array  maxabs(x - ref) <= RTOL * max(maxabs(ref), 1)   
RTOL = 1e-10


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_devices() -> int:
    """CUDA devices the C-ABI library sees (0 without a GPU / driver / built library)."""
    try:
        from rosdyn_b200 import _lib
        return int(_lib.load().rdb_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing them; on a GPU box nothing is skipped
    (and a missing library still fails loudly there: the tests themselves load it)."""
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (rdb_device_count() == 0)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def rel_err(x, ref):
    x, ref = np.asarray(x), np.asarray(ref)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    if x.size == 0:
        return 0.0
    return float(np.max(np.abs(x - ref)) / max(float(np.max(np.abs(ref))), 1.0))


def assert_close(x, ref, what="", rtol=RTOL):
    e = rel_err(x, ref)
    assert e <= rtol, f"{what}: rel err {e:.3e} > {rtol:.1e}"


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")))
    return load


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    from oracle import oracle
    oracle.build()
