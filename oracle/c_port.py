"""ORACLE, ctypes face of the C/OpenMP port (oracle/c/ns_oracle_c.c) -- test infrastructure and CPU
baseline only; the product never imports this.  PARITY UNPINNED (see oracle/assemble.py).

The C port restates the reference's assembly loops (src/classes/NavierStokes.cpp:569-831) and its
solve_linear_system (cpp:833-868, NavierStokes.hpp:279-366) with OpenMP, so that the CPU baseline of
bench.py uses every host core the way the reference's MPI ranks would."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libns_oracle_c.so")
_lib = None

_c_d = ctypes.POINTER(ctypes.c_double)


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "c")])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = ctypes.CDLL(LIB)
        _lib.nso_num_threads.restype = ctypes.c_int
        _lib.nso_solve.restype = ctypes.c_int
        _lib.nso_solve_blocks.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def num_threads():
    return lib().nso_num_threads()


def assemble_linearized_raw(dim, points, cells, cell_dofs, N, pattern, is_c, cval, sol_old, sol_old_old, dt, theta, nu,
                            use_supg, gamma, first_order_ustar, with_pressure_matrices=True):
    """Array-level entry: returns (A, b, Mp, Kp) aligned with pattern = (rowptr, col)."""
    rowptr, col = pattern
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    nnz = col.shape[0]
    A = np.empty(nnz)
    b = np.empty(N)
    Mp = np.empty(nnz) if with_pressure_matrices else None
    Kp = np.empty(nnz) if with_pressure_matrices else None
    is_c = np.ascontiguousarray(is_c, dtype=np.uint8)
    cval = np.ascontiguousarray(cval, dtype=np.float64)
    pts = np.ascontiguousarray(points[:, :dim], dtype=np.float64)
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    cdofs = np.ascontiguousarray(cell_dofs, dtype=np.int32)
    so = np.ascontiguousarray(sol_old, dtype=np.float64)
    soo = np.ascontiguousarray(sol_old_old, dtype=np.float64)
    lib().nso_assemble_linearized(
        ctypes.c_int(dim), ctypes.c_int64(cells.shape[0]), _p(pts), _p(cells), _p(cdofs), ctypes.c_int64(N),
        _p(rowptr), _p(col), _p(is_c), _p(cval), _p(so), _p(soo), ctypes.c_double(dt), ctypes.c_double(theta),
        ctypes.c_double(nu), ctypes.c_int(int(use_supg)), ctypes.c_double(gamma), ctypes.c_int(int(first_order_ustar)),
        _p(A), _p(b), _p(Mp), _p(Kp))
    return A, b, Mp, Kp


def assemble_newton(mesh, dm, pattern, p, con, sol_current, sol_old, with_pressure_matrices=True):
    """Same contract as oracle.assemble.assemble(kind='newton'); returns (A, b, Mp, Kp)."""
    rowptr, col = pattern
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    N, nnz = dm.n_dofs, col.shape[0]
    A, b = np.empty(nnz), np.empty(N)
    Mp = np.empty(nnz) if with_pressure_matrices else None
    Kp = np.empty(nnz) if with_pressure_matrices else None
    is_c = np.ascontiguousarray(con.is_c, dtype=np.uint8)
    cval = np.ascontiguousarray(con.val, dtype=np.float64)
    pts = np.ascontiguousarray(mesh.points[:, :mesh.dim], dtype=np.float64)
    cells = np.ascontiguousarray(mesh.cells, dtype=np.int32)
    cdofs = np.ascontiguousarray(dm.cell_dofs, dtype=np.int32)
    sc = np.ascontiguousarray(sol_current, dtype=np.float64)
    so = np.ascontiguousarray(sol_old, dtype=np.float64)
    lib().nso_assemble_newton(
        ctypes.c_int(mesh.dim), ctypes.c_int64(cells.shape[0]), _p(pts), _p(cells), _p(cdofs), ctypes.c_int64(N),
        _p(rowptr), _p(col), _p(is_c), _p(cval), _p(sc), _p(so), ctypes.c_double(p.dt), ctypes.c_double(p.theta),
        ctypes.c_double(p.nu), ctypes.c_int(int(p.use_supg)), ctypes.c_double(p.gamma), _p(A), _p(b), _p(Mp), _p(Kp))
    return A, b, Mp, Kp


def assemble_linearized(mesh, dm, pattern, p, con, sol_old, sol_old_old, with_pressure_matrices=True):
    """Same contract as oracle.assemble.assemble(kind='linearized'); returns (A, b, Mp, Kp)."""
    return assemble_linearized_raw(mesh.dim, mesh.points, mesh.cells, dm.cell_dofs, dm.n_dofs, pattern, con.is_c, con.val,
                                   sol_old, sol_old_old, p.dt, p.theta, p.nu, p.use_supg, p.gamma, p.first_order_ustar,
                                   with_pressure_matrices)


def solve(pattern, N, n_u, A, Mp, Kp, b, p, max_it=200, tol_rel=1e-2, n_tmp_vectors=150, nblocks=None, schur_mass_coeff=-1.0, kp_tol=1e-4):
    """solve_linear_system(): returns (x, iterations, residual, converged).  nblocks = number of ILU
    row blocks (the reference's MPI rank count); default = the OpenMP thread count."""
    rowptr, col = pattern
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    x = np.zeros(N)
    it = ctypes.c_int(0)
    res = ctypes.c_double(0)
    if nblocks is None:
        nblocks = num_threads()
    rc = lib().nso_solve(ctypes.c_int64(N), ctypes.c_int64(n_u), _p(rowptr), _p(col), _p(A), _p(Mp), _p(Kp), _p(b),
                         ctypes.c_double(p.nu), ctypes.c_double(p.rho), ctypes.c_double(p.dt), ctypes.c_double(p.theta),
                         ctypes.c_int(max_it), ctypes.c_double(tol_rel), ctypes.c_int(n_tmp_vectors), ctypes.c_int(nblocks), ctypes.c_double(schur_mass_coeff), ctypes.c_double(kp_tol),
                         _p(x), ctypes.byref(it), ctypes.byref(res))
    return x, it.value, res.value, rc == 0


def solve_blocks(pattern, N, n_u, A, pp, b, p, max_it=200, tol_rel=1e-2, n_tmp_vectors=150, nblocks=None, schur_mass_coeff=-1.0,
                 kp_tol=1e-4, time_budget_s=0.0):
    """solve_linear_system() with M_p / K_p given as the compact pressure-block CSR of oracle.assemble.pressure_blocks
    (pp = (ptr, col, Mp, Kp)); A is used in place.  Returns (x, iterations, residual, status, timings) with status
    0 converged / 1 max_it reached / 2 stopped by time_budget_s, timings = dict(setup, gmres, kp_cg, total) seconds."""
    rowptr, col = pattern
    assert rowptr.dtype == np.int64 and col.dtype == np.int32 and rowptr.flags.c_contiguous and col.flags.c_contiguous
    pptr, pcol, Mp, Kp = pp
    x = np.zeros(N)
    it = ctypes.c_int(0)
    res = ctypes.c_double(0)
    tm = np.zeros(4)
    if nblocks is None:
        nblocks = num_threads()
    rc = lib().nso_solve_blocks(ctypes.c_int64(N), ctypes.c_int64(n_u), _p(rowptr), _p(col), _p(A), _p(np.ascontiguousarray(pptr, np.int64)),
                                _p(np.ascontiguousarray(pcol, np.int32)), _p(np.ascontiguousarray(Mp)), _p(np.ascontiguousarray(Kp)), _p(b),
                                ctypes.c_double(p.nu), ctypes.c_double(p.rho), ctypes.c_double(p.dt), ctypes.c_double(p.theta),
                                ctypes.c_int(max_it), ctypes.c_double(tol_rel), ctypes.c_int(n_tmp_vectors), ctypes.c_int(nblocks),
                                ctypes.c_double(schur_mass_coeff), ctypes.c_double(kp_tol), ctypes.c_double(time_budget_s),
                                _p(x), ctypes.byref(it), ctypes.byref(res), _p(tm))
    return x, it.value, res.value, rc, dict(setup=tm[0], gmres=tm[1], kp_cg=tm[2], total=tm[3])


def set_num_threads(n):
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline must use the host cores regardless."""
    lib().nso_set_num_threads(ctypes.c_int(int(n)))
