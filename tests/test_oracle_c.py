"""The oracle's C/OpenMP port (oracle/c/ns_oracle_c.c, the CPU baseline of bench.py) against the numpy
oracle: assembly to round-off, ILU(k) against a dense textbook ILU(k), and the GMRES solve against a
direct solve.  CPU only."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import assemble as asm, c_port, dofs as odofs, postprocess as pp, solve as osolve
from tests.conftest import synthetic_state


def _problem(mesh, supg, first_step):
    dim = mesh.dim
    dm = odofs.enumerate_dofs(mesh)
    pat = odofs.make_sparsity(dm)
    con = odofs.build_constraints(mesh, dm, pp.inlet_profile(dim, 1.0, False, 4.0, 1.0), pp.boundary_ids(dim))
    un, unm1 = synthetic_state(dm, dim, 1.0)
    p = asm.Params(dt=0.01, theta=0.5, nu=1e-3, use_supg=supg, first_step=first_step)
    return dm, pat, con, un, unm1, p


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("supg,first_step", [(False, False), (True, True), (True, False)])
def test_c_assembly_matches_numpy_oracle_2d(golden_mesh, supg, first_step):
    mesh = golden_mesh("mesh-2D")
    dm, pat, con, un, unm1, p = _problem(mesh, supg, first_step)
    ref = asm.assemble(mesh, dm, pat, p, con, "linearized", un, unm1)
    A, b, Mp, Kp = c_port.assemble_linearized(mesh, dm, pat, p, con, un, unm1)
    assert _rel(A, ref.A) < 1e-13 and _rel(b, ref.b) < 1e-13
    assert _rel(Mp, ref.Mp) < 1e-13 and _rel(Kp, ref.Kp) < 1e-13


@pytest.mark.parametrize("supg", [False, True])
def test_c_newton_assembly_matches_numpy_oracle_2d(golden_mesh, supg):
    mesh = golden_mesh("mesh-2D")
    dm = odofs.enumerate_dofs(mesh)
    pat = odofs.make_sparsity(dm)
    con = odofs.build_constraints(mesh, dm, None, pp.boundary_ids(2), homogeneous=True)
    uk, un = synthetic_state(dm, 2, 1.0)
    uk[dm.n_u:] = np.random.default_rng(3).standard_normal(dm.n_p)
    p = asm.Params(dt=0.05, theta=1.0, nu=1e-3, use_supg=supg)
    ref = asm.assemble(mesh, dm, pat, p, con, "newton", uk, un)
    A, b, Mp, Kp = c_port.assemble_newton(mesh, dm, pat, p, con, uk, un)
    assert _rel(A, ref.A) < 1e-13 and _rel(b, ref.b) < 1e-13
    assert _rel(Mp, ref.Mp) < 1e-13 and _rel(Kp, ref.Kp) < 1e-13


def test_fast_sparsity_equals_reference_construction(golden_mesh, small_3d_mesh):
    for mesh in (golden_mesh("mesh-2D"), small_3d_mesh):
        dm = odofs.enumerate_dofs(mesh)
        a, b = odofs.make_sparsity(dm), odofs.make_sparsity_fast(dm)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_c_assembly_and_solve_3d(small_3d_mesh):
    mesh = small_3d_mesh
    dm, pat, con, un, unm1, p = _problem(mesh, True, False)
    ref = asm.assemble(mesh, dm, pat, p, con, "linearized", un, unm1)
    A, b, Mp, Kp = c_port.assemble_linearized(mesh, dm, pat, p, con, un, unm1)
    assert _rel(A, ref.A) < 1e-13 and _rel(b, ref.b) < 1e-13
    N = dm.n_dofs
    # Newton system with SUPG + grad-div in 3-D (cpp:278-539)
    uk = un.copy()
    uk[dm.n_u:] = np.random.default_rng(3).standard_normal(dm.n_p)
    conh = odofs.build_constraints(mesh, dm, None, pp.boundary_ids(3), homogeneous=True)
    refn = asm.assemble(mesh, dm, pat, p, conh, "newton", uk, unm1)
    An, bn, _, _ = c_port.assemble_newton(mesh, dm, pat, p, conh, uk, unm1)
    assert _rel(An, refn.A) < 1e-13 and _rel(bn, refn.b) < 1e-13
    x, its, res, ok = c_port.solve(pat, N, dm.n_u, A, Mp, Kp, b, p, nblocks=4)
    assert ok and 0 < its < 200
    Ac = asm.to_csr(pat, A, N)
    xd = osolve.direct_solve(Ac, b)
    assert _rel(x, xd) < 1e-2          # stopped at 1e-2 * ||b|| on the preconditioned residual
    # a tighter tolerance gets closer: the solver is consistent, not just terminating
    x2, its2, _, ok2 = c_port.solve(pat, N, dm.n_u, A, Mp, Kp, b, p, tol_rel=1e-8, max_it=500, nblocks=4, kp_tol=1e-12)
    assert ok2 and its2 > its and _rel(x2, xd) < 1e-6


def _dense_iluk(Ad, lof):
    n = Ad.shape[0]
    big = 10 ** 6
    lev = np.where(Ad != 0, 0, big)
    np.fill_diagonal(lev, 0)
    LU = Ad.copy()
    for i in range(n):
        for k in range(i):
            if lev[i, k] > lof:
                continue
            LU[i, k] /= LU[k, k]
            for j in range(k + 1, n):
                if lev[k, j] > lof:
                    continue
                nl = lev[i, k] + lev[k, j] + 1
                if lev[i, j] > lof and nl > lof:
                    continue
                if lev[i, j] > lof:
                    LU[i, j] = 0.0
                lev[i, j] = min(lev[i, j], nl)
                LU[i, j] -= LU[i, k] * LU[k, j]
        LU[i, lev[i] > lof] = 0
    return LU, int((lev <= lof).sum())


@pytest.mark.parametrize("lof", [0, 1, 2])
def test_c_iluk_matches_dense_textbook(lof):
    n = 150
    M = sp.csr_matrix(sp.random(n, n, density=0.03, random_state=3, format="csr") + 3 * sp.eye(n))
    M.sort_indices()
    x = np.random.default_rng(1).standard_normal(n)
    rhs = M @ x
    LU, nnz_ref = _dense_iluk(M.toarray(), lof)
    y_ref = np.linalg.solve(np.triu(LU), np.linalg.solve(np.tril(LU, -1) + np.eye(n), rhs))
    ptr, col, val = M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data.copy()
    y = np.zeros(n)
    nnz = ctypes.c_int64()
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    c_port.lib().nso_ilu_apply_once(ctypes.c_int(n), P(ptr), P(col), P(val), ctypes.c_int(lof), ctypes.c_int(1), P(rhs), P(y),
                                    ctypes.byref(nnz))
    assert nnz.value == nnz_ref
    assert _rel(y, y_ref) < 1e-12


@pytest.mark.parametrize("case", ["2D-2", "2D-1"])
def test_oracle_trajectory_same_with_c_assembly(golden_mesh, case):
    """Oracle(c_assembly=True) -- used for the long known-answer runs -- follows the numpy oracle step by step."""
    mesh = golden_mesh("mesh-2D")
    a = osolve.Oracle(mesh, case, solver="direct")
    b = osolve.Oracle(mesh, case, solver="direct", c_assembly=True)
    for _ in range(3):
        ia, ib = a.step(), b.step()
        for k in ("cd", "cl", "dp"):
            assert abs(ia[k] - ib[k]) <= 1e-9 * max(1.0, abs(ia[k]))
    assert np.linalg.norm(a.current_solution - b.current_solution) <= 1e-10 * np.linalg.norm(a.current_solution)
