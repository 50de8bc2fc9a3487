import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_DIR = os.path.join(ROOT, "navier-stokes_equations_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_nsb():
    """The product package (navier-stokes_equations_b200/, loaded by path through the root module nsb200.py)."""
    import nsb200
    return nsb200


@pytest.fixture(scope="session")
def nsb():
    return load_nsb()


@pytest.fixture(scope="session")
def golden_mesh():
    from tools import msh
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = msh.load_npz(os.path.join(GOLDEN, name + ".npz"))
        return cache[name]

    return get


@pytest.fixture(scope="session")
def msh_file(golden_mesh, tmp_path_factory):
    """Writes a golden mesh as MSH 2.2 ASCII (what the reference reads) and returns the path."""
    from tools import msh
    d = tmp_path_factory.mktemp("meshes")
    cache = {}

    def get(name):
        if name not in cache:
            path = str(d / (name + ".msh"))
            msh.write_msh(path, golden_mesh(name))
            cache[name] = path
        return cache[name]

    return get


@pytest.fixture(scope="session")
def small_3d_mesh():
    from tools import meshgen
    return meshgen.mesh_3d(lc_cyl=0.05, lc_global=0.15)


def synthetic_state(dm, dim, U_m):
    """inlet paraboloid times (1 + 0.1 xi), xi ~ U(-1,1), seeds 1234 / 1235."""
    from oracle import postprocess as pp
    N = dm.n_dofs
    full = pp.inlet_profile(dim, U_m, False, 0.0, 0.0)
    base = np.zeros(N)
    base[:dm.n_u] = full(dm.support_points[:dm.n_u], dm.component[:dm.n_u])
    un = base * (1 + 0.1 * np.random.default_rng(1234).uniform(-1, 1, N))
    unm1 = base * (1 + 0.1 * np.random.default_rng(1235).uniform(-1, 1, N))
    return un, unm1
