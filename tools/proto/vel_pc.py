"""CPU prototype (research script, not product, not a test): preconditioners for the velocity block F of the linearised
3D-2Z system, on a mesh-3D-5-equivalent with parameters rescaled so that the non-dimensional groups match the
mesh-3D-20-equivalent (h x4  =>  dt x16, U /4; gamma dt / h^2, nu dt / h^2 and U dt / h unchanged)."""
import sys, time, os
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import assemble as asm, dofs as odofs, postprocess as pp, c_port, fe_tables as fe
from oracle.solve import gmres_left, NoConvergence
from tools import meshgen

LINES3 = [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]


def build(level=5, scale=4.0, lc=None):
    mesh = meshgen.mesh_3d(level) if lc is None else meshgen.mesh_3d(lc_cyl=lc[0], lc_global=lc[1])
    dm = odofs.enumerate_dofs(mesh)
    rp, col = odofs.make_sparsity_fast(dm)
    pat = (np.ascontiguousarray(rp, np.int64), np.ascontiguousarray(col, np.int32))
    U_m = 2.25 / scale
    dt = 0.01 * scale * scale
    con = odofs.build_constraints(mesh, dm, pp.inlet_profile(3, U_m, False, 4.0, 1.0), pp.boundary_ids(3))
    N = dm.n_dofs
    H = 0.41
    pts = dm.support_points
    prof = 16.0 * U_m * pts[:, 0] * pts[:, 1] * (H - pts[:, 0]) * (H - pts[:, 1]) / H ** 4
    base = np.where((dm.component == 2) & (np.arange(N) < dm.n_u), prof, 0.0)
    un = base * (1 + 0.1 * np.random.default_rng(1234).uniform(-1, 1, N))
    unm1 = base * (1 + 0.1 * np.random.default_rng(1235).uniform(-1, 1, N))
    p = asm.Params(dt=dt, theta=0.5, nu=1e-3, use_supg=True)
    A, b, _, _ = c_port.assemble_linearized(mesh, dm, pat, p, con, un, unm1, with_pressure_matrices=False)
    A = sp.csr_matrix((A, pat[1], pat[0]), shape=(N, N))
    pptr, pcol, Mp, Kp = asm.pressure_blocks(mesh, dm, con)
    Mp = sp.csr_matrix((Mp, pcol, pptr), shape=(dm.n_p, dm.n_p))
    Kp = sp.csr_matrix((Kp, pcol, pptr), shape=(dm.n_p, dm.n_p))
    return dict(mesh=mesh, dm=dm, con=con, p=p, A=A, b=b, Mp=Mp, Kp=Kp)


def block_jacobi(F, bs=3):
    n = F.shape[0] // bs
    Fb = F.tobsr(blocksize=(bs, bs))
    Fb.sort_indices()
    D = np.zeros((n, bs, bs))
    for i in range(n):
        s, e = Fb.indptr[i], Fb.indptr[i + 1]
        k = s + np.searchsorted(Fb.indices[s:e], i)
        D[i] = Fb.data[k]
    Dinv = np.linalg.inv(D)
    return sp.bsr_matrix((Dinv, np.arange(n), np.arange(n + 1)), shape=F.shape).tocsr()


def prolongation(S):
    """P1-vector -> P2-vector prolongation in the velocity numbering (n_u x 3V)."""
    mesh, dm = S["mesh"], S["dm"]
    cd = dm.cell_dofs.astype(np.int64)
    cells = mesh.cells.astype(np.int64)
    rows, cols, vals = [], [], []
    for v in range(4):
        for c in range(3):
            rows.append(cd[:, v * 4 + c]); cols.append(3 * cells[:, v] + c); vals.append(np.ones(len(cd)))
    for l, (i, j) in enumerate(LINES3):
        for c in range(3):
            r = cd[:, 16 + 3 * l + c]
            rows += [r, r]; cols += [3 * cells[:, i] + c, 3 * cells[:, j] + c]; vals += [np.full(len(cd), 0.5)] * 2
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    key = rows * (3 * mesh.n_vertices) + cols
    _, idx = np.unique(key, return_index=True)
    P = sp.csr_matrix((vals[idx], (rows[idx], cols[idx])), shape=(dm.n_u, 3 * mesh.n_vertices))
    return P


class Cheb:
    """x ~ A^-1 b by `deg` steps of Chebyshev on Dinv A with spectrum in [lo, hi] (zero initial guess unless x0)."""
    def __init__(self, A, Dinv, lo, hi, deg):
        self.A, self.Dinv, self.lo, self.hi, self.deg = A, Dinv, lo, hi, deg
        self.nmv = 0

    def __call__(self, b, x0=None):
        A, Dinv = self.A, self.Dinv
        theta, delta = 0.5 * (self.hi + self.lo), 0.5 * (self.hi - self.lo)
        sigma = theta / delta
        rho = 1.0 / sigma
        if x0 is None:
            x = np.zeros_like(b); r = b.copy()
        else:
            x = x0.copy(); r = b - A @ x; self.nmv += 1
        d = (Dinv @ r) / theta
        for k in range(self.deg):
            x = x + d
            if k + 1 == self.deg:
                break
            r = r - A @ d; self.nmv += 1
            rho_new = 1.0 / (2.0 * sigma - rho)
            d = rho_new * rho * d + (2.0 * rho_new / delta) * (Dinv @ r)
            rho = rho_new
        return x


def lam_max(A, Dinv, its=30, seed=0):
    v = np.random.default_rng(seed).standard_normal(A.shape[0])
    for _ in range(its):
        w = Dinv @ (A @ v)
        l = np.linalg.norm(w) / np.linalg.norm(v)
        v = w / np.linalg.norm(w)
    return l


def outer(S, Finv, cM=None, tol=1e-2, max_it=200, verbose=True):
    """Left-preconditioned GMRES with the block-triangular operator (exact K_p, M_p solves)."""
    A, b, dm, p = S["A"], S["b"], S["dm"], S["p"]
    n_u = dm.n_u
    B = A[n_u:, :n_u].tocsr()
    if "Kinv" not in S:
        S["Kinv"] = spla.splu(S["Kp"].tocsc()).solve
        S["Minv"] = spla.splu(S["Mp"].tocsc()).solve
    cM = (p.theta * p.nu + p.gamma) if cM is None else cM

    def P(x):
        y0 = Finv(x[:n_u])
        t = x[n_u:] - B @ y0
        y1 = -(p.rho / p.dt) * S["Kinv"](t) - cM * S["Minv"](t)
        return np.concatenate([y0, y1])
    try:
        x, it, res = gmres_left(lambda v: A @ v, b, P, tol * np.linalg.norm(b), max_it)
        ok = True
    except NoConvergence as e:
        it, ok = e.last_step, False
    return it, ok


class TwoLevel:
    """Multiplicative two-level cycle: Chebyshev pre-smoothing (zero guess), coarse correction in the P1 space, Chebyshev
    post-smoothing.  coarse = callable r_c -> e_c.  Fixed linear operator."""
    def __init__(self, F, Dinv, P, coarse, lo, hi, pre, post):
        self.F, self.P, self.coarse = F, P, coarse
        self.pre = Cheb(F, Dinv, lo, hi, pre) if pre > 0 else None
        self.post = Cheb(F, Dinv, lo, hi, post) if post > 0 else None
        self.nmv = 0

    def __call__(self, b):
        F, P = self.F, self.P
        if self.pre is not None:
            x = self.pre(b)
            r = b - F @ x; self.nmv += 1
        else:
            x = np.zeros_like(b); r = b
        x = x + P @ self.coarse(P.T @ r)
        if self.post is not None:
            x = self.post(b, x0=x)
        return x


class Cycle:
    """General fixed cycle: list of ops applied in order to solve F x = b from x = 0:
       ("c",) coarse correction, ("s", deg, lo, hi) Chebyshev smoothing steps."""
    def __init__(self, F, Dinv, P, coarse, ops):
        self.F, self.Dinv, self.P, self.coarse, self.ops = F, Dinv, P, coarse, ops
        self.nmv = 0
        self.ncoarse = 0

    def __call__(self, b):
        F, P = self.F, self.P
        x = None
        for op in self.ops:
            if op[0] == "c":
                if x is None:
                    r = b
                else:
                    r = b - F @ x; self.nmv += 1
                e = P @ self.coarse(P.T @ r); self.ncoarse += 1
                x = e if x is None else x + e
            else:
                ch = Cheb(F, self.Dinv, op[2], op[3], op[1])
                x = ch(b, x0=x)
                self.nmv += ch.nmv
        return x


class PatchSchwarz:
    """Additive Schwarz over vertex stars: patch(v) = velocity DoFs of vertex v and of the midpoints of all edges at v
    (unconstrained ones).  z = sum_v R_v^T F_v^-1 R_v r.  Acts like `Dinv @ r`."""
    def __init__(self, S, F, Fpatch=None):
        mesh, dm = S["mesh"], S["dm"]
        Fp = F if Fpatch is None else Fpatch
        cd = dm.cell_dofs.astype(np.int64)
        cells = mesh.cells.astype(np.int64)
        isc = S["con"].is_c[:dm.n_u]
        V_ = mesh.n_vertices
        # node (first dof / 3) lists per vertex
        pairs = []
        for v in range(4):
            pairs.append(np.stack([cells[:, v], cd[:, v * 4] // 3], 1))
        for l, (i, j) in enumerate(LINES3):
            nd = cd[:, 16 + 3 * l] // 3
            pairs.append(np.stack([cells[:, i], nd], 1)); pairs.append(np.stack([cells[:, j], nd], 1))
        pr = np.unique(np.concatenate(pairs), axis=0)
        order = np.argsort(pr[:, 0], kind="stable")
        pr = pr[order]
        starts = np.searchsorted(pr[:, 0], np.arange(V_ + 1))
        self.idx, self.inv = [], []
        Fc = Fp.tocsr()
        self.n = F.shape[0]
        sizes = []
        for v in range(V_):
            nodes = pr[starts[v]:starts[v + 1], 1]
            dofs = (3 * nodes[:, None] + np.arange(3)[None, :]).ravel()
            dofs = dofs[~isc[dofs]]
            if dofs.size == 0:
                continue
            A = Fc[dofs][:, dofs].toarray()
            self.idx.append(dofs); self.inv.append(np.linalg.inv(A)); sizes.append(dofs.size)
        self.sizes = np.array(sizes)
        # constrained dofs: plain diagonal
        self.cdofs = np.where(isc)[0]
        self.cdiag = 1.0 / Fc.diagonal()[self.cdofs]
        # weights: number of patches containing each dof (for optional weighting)
        cnt = np.zeros(self.n)
        for d in self.idx:
            cnt[d] += 1
        self.cnt = np.maximum(cnt, 1)

    def __matmul__(self, r):
        z = np.zeros_like(r)
        for d, Ai in zip(self.idx, self.inv):
            z[d] += Ai @ r[d]
        z[self.cdofs] = self.cdiag * r[self.cdofs]
        return z
