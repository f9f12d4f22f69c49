import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def fast5_files():
    import glob
    fs = sorted(glob.glob(os.path.join(GOLDEN, "fast5", "*.fast5")))
    assert len(fs) == 5
    return fs


@pytest.fixture(scope="session")
def seg_golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "segmentation.npz"))


@pytest.fixture(scope="session")
def weights_by_species():
    from nanoreviser_b200 import weights
    cache = {}

    def get(sp):
        if sp not in cache:
            cache[sp] = weights.load_species(sp, os.path.join(ROOT, "model"))
        return cache[sp]
    return get


@pytest.fixture(scope="session")
def reads(fast5_files):
    from nanoreviser_b200 import fast5
    return [fast5.read_fast5_arrays(f) for f in fast5_files]


@pytest.fixture(scope="session")
def reviser_by_species(weights_by_species):
    """One CUDA handle per species, shared by all GPU tests.  Built library required: fails loudly."""
    from nanoreviser_b200 import engine
    cache = {}

    def get(sp):
        if sp not in cache:
            m1, m2 = weights_by_species(sp)
            cache[sp] = engine.Reviser(m1, m2, device=0)
        return cache[sp]
    yield get
    for r in cache.values():
        r.close()
