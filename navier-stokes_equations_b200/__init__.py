"""ctypes bindings of libnsb200.so (include/nsb200.h) and libnsbhost.so (include/nsb200_host.h).

Thin bindings only: the product is the CUDA/C++ behind the C ABI.  There is no CPU
fallback -- loading fails loudly if the shared library is missing, and every call raises
`NsbError` with the library's own message on a non-zero/negative return code.

The directory name contains a '-', so import it by path (tests/conftest.py: `load_nsb()`).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnsb200.so")

NSB_SOLUTION_OLD, NSB_SOLUTION_OLD_OLD, NSB_CURRENT_SOLUTION, NSB_SOLUTION, NSB_RHS = 0, 1, 2, 3, 4
PROFILE_CLASSES = ("asm_context", "asm_rows", "spmv", "spmv_vel", "schur", "amg", "orth", "other")


class NsbError(RuntimeError):
    pass


class NsbParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("theta", C.c_double), ("nu", C.c_double), ("rho", C.c_double),
                ("gamma", C.c_double), ("use_supg", C.c_int32), ("first_order_ustar", C.c_int32)]


class NsbSolverOpts(C.Structure):
    _fields_ = [("poly_degree_F", C.c_int32), ("poly_refresh", C.c_int32), ("poly_target", C.c_double), ("cheb_degree_Mp", C.c_int32),
                ("amg_smoother_degree", C.c_int32), ("schur_mass_coeff", C.c_double), ("reorthogonalize", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NsbError("libnsb200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.nsb_last_error.restype = C.c_char_p
    return _lib


def _p(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class Device:
    """One nsb_handle.  Methods mirror include/nsb200.h one to one."""

    def __init__(self, dim, device=0):
        self.dim = dim
        self.h = C.c_void_p()
        L = lib()
        rc = L.nsb_create(dim, device, C.byref(self.h))
        if rc != 0:
            msg = L.nsb_last_error(self.h).decode() if self.h else "nsb_create failed"
            if self.h:
                L.nsb_destroy(self.h)
                self.h = None
            raise NsbError(msg)
        self.n_dofs = 0

    def _ck(self, rc, soft=()):
        if rc != 0 and rc not in soft:
            raise NsbError(lib().nsb_last_error(self.h).decode())
        return rc

    def close(self):
        if self.h:
            lib().nsb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup
    def comm_init(self, rank, nranks, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._ck(lib().nsb_comm_init(self.h, rank, nranks, buf))

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        if lib().nsb_comm_unique_id(buf) != 0:
            raise NsbError("ncclGetUniqueId failed")
        return buf.raw

    def upload_mesh(self, coords, cell_vertices, cell_dofs, n_u, n_p, cell_part=None):
        coords = np.ascontiguousarray(coords, np.float64)
        cv = np.ascontiguousarray(cell_vertices, np.uint32)
        cd = np.ascontiguousarray(cell_dofs, np.uint32)
        part = None if cell_part is None else np.ascontiguousarray(cell_part, np.int32)
        self.n_dofs = int(n_u + n_p)
        self._ck(lib().nsb_upload_mesh(self.h, C.c_int64(coords.shape[0]), _p(coords, C.c_double),
                                        C.c_int64(cv.shape[0]), _p(cv, C.c_uint32), _p(cd, C.c_uint32),
                                        C.c_int64(n_u), C.c_int64(n_p),
                                        _p(part, C.c_int32) if part is not None else None))

    def sizes(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(lib().nsb_get_sizes(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def pattern(self):
        n, nnz, _ = self.sizes()
        rp = np.empty(n + 1, np.int64)
        col = np.empty(nnz, np.uint32)
        self._ck(lib().nsb_get_pattern(self.h, _p(rp, C.c_int64), _p(col, C.c_uint32)))
        return rp, col

    def row_gids(self):
        n, _, _ = self.sizes()
        g = np.empty(n, np.int64)
        self._ck(lib().nsb_get_row_gids(self.h, _p(g, C.c_int64)))
        return g

    # ---- per-step inputs
    def set_constraints(self, dofs, vals):
        d = np.ascontiguousarray(dofs, np.uint32)
        v = np.ascontiguousarray(vals, np.float64)
        self._ck(lib().nsb_set_constraints(self.h, C.c_int64(d.shape[0]), _p(d, C.c_uint32), _p(v, C.c_double)))

    def set_params(self, dt, theta, nu, rho=1.0, gamma=0.1, use_supg=False, first_order_ustar=False):
        p = NsbParams(dt, theta, nu, rho, gamma, int(use_supg), int(first_order_ustar))
        self._ck(lib().nsb_set_params(self.h, C.byref(p)))

    def set_solver_opts(self, poly_degree_F=0, poly_refresh=0, poly_target=0.0, cheb_degree_Mp=0, amg_smoother_degree=0,
                        schur_mass_coeff=0.0, reorthogonalize=1):
        o = NsbSolverOpts(poly_degree_F, poly_refresh, poly_target, cheb_degree_Mp, amg_smoother_degree, schur_mass_coeff, reorthogonalize)
        self._ck(lib().nsb_set_solver_opts(self.h, C.byref(o)))

    def set_vector(self, which, v):
        v = np.ascontiguousarray(v, np.float64)
        assert v.shape[0] == self.n_dofs
        self._ck(lib().nsb_set_vector(self.h, which, _p(v, C.c_double)))

    def get_vector(self, which, out=None):
        out = np.zeros(self.n_dofs) if out is None else out
        self._ck(lib().nsb_get_vector(self.h, which, _p(out, C.c_double)))
        return out

    def copy_vector(self, dst, src):
        self._ck(lib().nsb_copy_vector(self.h, dst, src))

    def axpy_vector(self, dst, alpha, src):
        self._ck(lib().nsb_axpy_vector(self.h, dst, C.c_double(alpha), src))

    # ---- hot path
    def assemble_linearized(self):
        self._ck(lib().nsb_assemble_linearized(self.h))

    def assemble_newton(self):
        self._ck(lib().nsb_assemble_newton(self.h))

    def assemble_pressure_matrices(self):
        self._ck(lib().nsb_assemble_pressure_matrices(self.h))

    def rhs_norm(self):
        v = C.c_double()
        self._ck(lib().nsb_rhs_norm(self.h, C.byref(v)))
        return v.value

    def solve(self, max_it=200, tol_rel=1e-2, n_tmp_vectors=150):
        it, res = C.c_int(), C.c_double()
        rc = self._ck(lib().nsb_solve(self.h, max_it, C.c_double(tol_rel), n_tmp_vectors, C.byref(it), C.byref(res)), soft=(1,))
        return rc == 0, it.value, res.value

    # ---- inspection
    def matrix_values(self):
        _, nnz, _ = self.sizes()
        v = np.empty(nnz)
        self._ck(lib().nsb_get_matrix_values(self.h, _p(v, C.c_double)))
        return v

    def pressure_matrix(self, which):
        n, nnz = C.c_int64(), C.c_int64()
        self._ck(lib().nsb_get_pressure_matrix(self.h, which, C.byref(n), C.byref(nnz), None, None, None))
        rp = np.empty(n.value + 1, np.int32)
        col = np.empty(nnz.value, np.int32)
        val = np.empty(nnz.value)
        self._ck(lib().nsb_get_pressure_matrix(self.h, which, C.byref(n), C.byref(nnz), _p(rp, C.c_int32), _p(col, C.c_int32), _p(val, C.c_double)))
        return rp, col, val

    def spmv(self, x):
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros(self.n_dofs)
        self._ck(lib().nsb_spmv(self.h, _p(x, C.c_double), _p(y, C.c_double)))
        return y

    # ---- measurement
    def timer_start(self):
        self._ck(lib().nsb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._ck(lib().nsb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def synchronize(self):
        self._ck(lib().nsb_synchronize(self.h))

    def profile_enable(self, on=True):
        self._ck(lib().nsb_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._ck(lib().nsb_profile_reset(self.h))

    def profile(self):
        out = {}
        for name in PROFILE_CLASSES:
            ms, n = C.c_double(), C.c_int64()
            self._ck(lib().nsb_profile_get(self.h, name.encode(), C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    def solver_info(self):
        d, r, l = C.c_int(), C.c_double(), C.c_int()
        self._ck(lib().nsb_solver_info(self.h, C.byref(d), C.byref(r), C.byref(l)))
        return dict(poly_degree=d.value, poly_probe_residual=r.value, amg_levels=l.value)

    def launch_count(self):
        n = C.c_int64()
        self._ck(lib().nsb_launch_count(self.h, C.byref(n)))
        return n.value
