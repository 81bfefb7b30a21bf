import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The suite needs dr-mpc_b200/libdre.so (built artefacts are not in git). The driver builds it with
    __graft_entry__.build() before the tests; a fresh clone gets the same build here (nvcc cross-compiles without a GPU).
    This builds the PRODUCT, it is not a fallback: without nvcc the library stays missing and the tests that load it fail loudly."""
    so = os.path.join(REPO, "dr-mpc_b200", "libdre.so")
    if not os.path.exists(so):
        import shutil
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            import __graft_entry__ as g
            g.build()
    yield


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def golden_orca():
    return np.load(os.path.join(GOLDEN, "orca_crowds.npz"))


@pytest.fixture(scope="session")
def golden_mpc():
    return np.load(os.path.join(GOLDEN, "mpc_sequences.npz"))


@pytest.fixture(scope="session")
def golden_env():
    return np.load(os.path.join(GOLDEN, "env_rollouts.npz"))


ORCA_KEYS = ["n1_s2.0", "n6_s3.0", "n6_s1.2", "n10_s4.0", "n10_s1.8", "n20_s5.0", "n20_s2.5", "n50_s8.0", "n50_s4.5"]


def mpc_path_table(g):
    paths = [g[f"path{i}"] for i in range(4)]
    stride = max(p.shape[1] for p in paths)
    tab = np.zeros((4, 3, stride))
    lens = np.zeros(4, dtype=np.int32)
    for i, p in enumerate(paths):
        tab[i, :, :p.shape[1]] = p
        lens[i] = p.shape[1]
    return paths, tab, lens
