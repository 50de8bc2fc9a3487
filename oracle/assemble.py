"""ORACLE (test infrastructure, not product code) -- cell-wise assembly.

PARITY UNPINNED by the reference (it has no tests / golden vectors and cannot be built
here: deal.II + Trilinos + MPI are absent).  The one external known answer the oracle is held
to is the published DFG benchmark 2D-1 (tests/test_oracle_pins.py::test_dfg_2d1_known_answer).
This is a CPU restatement of

  NavierStokes<dim>::assemble_linearized_system()   reference src/classes/NavierStokes.cpp:569-831
  NavierStokes<dim>::assemble_newton_system()       reference src/classes/NavierStokes.cpp:278-539
  AffineConstraints::distribute_local_to_global     call sites cpp:516-523, 810-817

written with the reference's own per-(i,j,q) formulas on vector-valued shape functions
(no block-structure shortcuts -- the CUDA path uses those, this file is the checker).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import it.
"""
import numpy as np
import scipy.sparse as sp
from . import fe_tables as fe


class Params:
    def __init__(self, dt, theta, nu, rho=1.0, use_supg=False, gamma=0.1,
                 first_step=False, second_step=False, backward_euler=False):
        self.dt, self.theta, self.nu, self.rho = dt, theta, nu, rho
        self.use_supg, self.gamma = use_supg, gamma
        # u* = u^n when first_step || second_step || time_scheme == BackwardEuler (cpp:665)
        self.first_order_ustar = bool(first_step or second_step or backward_euler)


class CellGeometry:
    """FEValues::reinit for the affine P1 mapping."""

    def __init__(self, mesh):
        dim = mesh.dim
        X = mesh.points[mesh.cells]                    # (C, nv, dim)
        J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))   # columns x_k - x_0
        self.detJ = np.linalg.det(J)
        if np.any(self.detJ <= 0):
            raise ValueError("mesh has cells with non-positive measure")
        Jinv = np.linalg.inv(J)
        # grad lambda_k = row k-1 of J^{-1} (k>=1), grad lambda_0 = -sum
        gl = np.empty((X.shape[0], dim + 1, dim))
        gl[:, 1:, :] = Jinv
        gl[:, 0, :] = -Jinv.sum(axis=1)
        self.grad_lambda = gl
        self.X = X
        # cell->diameter(): longest vertex distance
        d = X[:, :, None, :] - X[:, None, :, :]
        self.h = np.sqrt((d ** 2).sum(-1)).max(axis=(1, 2))
        self.dim = dim


def _basis_at(dim, geom, sl, lam_q):
    """Shape values / gradients of all dofs_per_cell vector-valued functions at ONE
    quadrature point, for the cells in slice sl.
      phi_u (K, dim)        grad_phi_u (c, K, dim, dim)  [comp][deriv]
      div_phi_u (c, K)      phi_p (K,)                   grad_phi_p (c, K, dim)"""
    node, comp = fe.local_dof_layout(dim)
    K = node.shape[0]
    N, dN = fe.p2_values(dim, lam_q[None, :])
    N, dN = N[0], dN[0]                                      # (nn,), (nn, nv)
    gl = geom.grad_lambda[sl]                                # (c, nv, dim)
    gradN = np.einsum("nk,ckd->cnd", dN, gl)                 # (c, nn, dim)
    c = gl.shape[0]
    phi_u = np.zeros((K, dim))
    grad_phi_u = np.zeros((c, K, dim, dim))
    div_phi_u = np.zeros((c, K))
    phi_p = np.zeros(K)
    grad_phi_p = np.zeros((c, K, dim))
    for k in range(K):
        if comp[k] < dim:
            phi_u[k, comp[k]] = N[node[k]]
            grad_phi_u[:, k, comp[k], :] = gradN[:, node[k], :]
            div_phi_u[:, k] = gradN[:, node[k], comp[k]]
        else:
            phi_p[k] = lam_q[node[k]]
            grad_phi_p[:, k, :] = gl[:, node[k], :]
    return phi_u, grad_phi_u, div_phi_u, phi_p, grad_phi_p


def _function_values(dim, phi_u, grad_phi_u, vec_local):
    """get_function_values / get_function_gradients of the velocity part."""
    val = np.einsum("ck,kd->cd", vec_local, phi_u)
    grad = np.einsum("ck,ckde->cde", vec_local, grad_phi_u)
    return val, grad


def _tau(p, u_mag, h):
    # tau = ((2/dt)^2 + (2|u|/h)^2 + (4 nu/h^2)^2)^(-1/2)      cpp:727-729 / 445-448
    return 1.0 / np.sqrt((2.0 / p.dt) ** 2 + (2.0 * u_mag / h) ** 2 + (4.0 * p.nu / (h * h)) ** 2)


def cell_matrices_linearized(mesh, dm, p: Params, sol_old, sol_old_old, sl, forcing=None, geom=None):
    """Local matrices / rhs of the linearised system for cells[sl] (cpp:642-805).
    Returns (cell_matrix (c,K,K), cell_rhs (c,K), cell_Mp (c,K,K), cell_Kp (c,K,K))."""
    dim = mesh.dim
    geom = geom or CellGeometry(mesh)
    pts, w = fe.quadrature(dim)
    lam = fe.barycentric(pts)
    cd = dm.cell_dofs[sl]
    c, K = cd.shape
    uo = sol_old[cd]
    uoo = sol_old_old[cd]
    h = geom.h[sl]
    A = np.zeros((c, K, K))
    b = np.zeros((c, K))
    Mp = np.zeros((c, K, K))
    Kp = np.zeros((c, K, K))
    dt, th, nu = p.dt, p.theta, p.nu
    for q in range(len(w)):
        JxW = w[q] * geom.detJ[sl]
        phi_u, G, div, phi_p, gpp = _basis_at(dim, geom, sl, lam[q])
        u_old, gu_old = _function_values(dim, phi_u, G, uo)
        u_oo, _ = _function_values(dim, phi_u, G, uoo)
        if p.first_order_ustar:
            u_star = u_old.copy()
        else:
            u_star = 2.0 * u_old - u_oo
            ns = np.sqrt((u_star ** 2).sum(1))
            no = np.sqrt((u_old ** 2).sum(1))
            clamp = (no > 1e-12) & (ns > 1.2 * no)
            u_star[clamp] = u_old[clamp]
        f_new = np.zeros((c, dim)) if forcing is None else forcing[0][sl, q]
        f_old = np.zeros((c, dim)) if forcing is None else forcing[1][sl, q]
        f_th = th * f_new + (1.0 - th) * f_old
        # ---- rhs (cpp:702-720)
        rhs_mass = (1.0 / dt) * np.einsum("cd,kd->ck", u_old, phi_u)
        rhs_visc = -(1.0 - th) * nu * np.einsum("cde,ckde->ck", gu_old, G)
        conv_old = np.einsum("cde,ce->cd", gu_old, u_old)
        rhs_conv = -(1.0 - th) * np.einsum("cd,kd->ck", conv_old, phi_u)
        b += (rhs_mass + rhs_visc + rhs_conv) * JxW[:, None]
        b += np.einsum("cd,kd->ck", f_th, phi_u) * JxW[:, None]
        # G_j u*  -> (c,K,dim)
        Gu = np.einsum("ckde,ce->ckd", G, u_star)
        if p.use_supg:
            tau = _tau(p, np.sqrt((u_star ** 2).sum(1)), h)
            # rhs test vector: tau * (u_star * grad_phi_u[i])  -- contraction over the FIRST index (cpp:733)
            supg_rhs = tau[:, None, None] * np.einsum("cd,ckde->cke", u_star, G)
            rhs_source = f_th + u_old / dt
            b += np.einsum("cke,ce->ck", supg_rhs, rhs_source) * JxW[:, None]
        # ---- matrix (cpp:748-795)
        val = (1.0 / dt) * np.einsum("id,jd->ij", phi_u, phi_u)[None]
        val = val + th * nu * np.einsum("cide,cjde->cij", G, G)
        val = val + th * np.einsum("cjd,id->cij", Gu, phi_u)
        val = val - np.einsum("j,ci->cij", phi_p, div)
        val = val - np.einsum("i,cj->cij", phi_p, div)
        A += val * JxW[:, None, None]
        if p.use_supg:
            supg = tau[:, None, None] * Gu                               # tau * (grad_phi_u[i] * u_star), cpp:774
            op = phi_u[None] / dt + Gu                                   # phi_j/dt + grad_phi_u[j]*u_star
            A += np.einsum("cid,cjd->cij", supg, op) * JxW[:, None, None]
            A += np.einsum("cid,cjd->cij", supg, gpp) * JxW[:, None, None]
            A += p.gamma * np.einsum("ci,cj->cij", div, div) * JxW[:, None, None]
        Mp += np.einsum("i,j->ij", phi_p, phi_p)[None] * JxW[:, None, None]
        Kp += np.einsum("cid,cjd->cij", gpp, gpp) * JxW[:, None, None]
    return A, b, Mp, Kp


def cell_matrices_newton(mesh, dm, p: Params, sol_cur, sol_old, sl, forcing=None, geom=None):
    """Local Jacobian / -residual of the Newton system for cells[sl] (cpp:327-512)."""
    dim = mesh.dim
    geom = geom or CellGeometry(mesh)
    pts, w = fe.quadrature(dim)
    lam = fe.barycentric(pts)
    node, comp = fe.local_dof_layout(dim)
    H2 = fe.p2_second(dim)
    cd = dm.cell_dofs[sl]
    c, K = cd.shape
    uc = sol_cur[cd]
    uo = sol_old[cd]
    h = geom.h[sl]
    gl = geom.grad_lambda[sl]
    A = np.zeros((c, K, K))
    b = np.zeros((c, K))
    Mp = np.zeros((c, K, K))
    Kp = np.zeros((c, K, K))
    dt, th, nu = p.dt, p.theta, p.nu
    # laplacian of each scalar P2 node function: sum_d d2N/dx_d^2 = sum_kl H_kl (grad l_k . grad l_l)
    glgl = np.einsum("ckd,cld->ckl", gl, gl)
    lapN = np.einsum("nkl,ckl->cn", H2, glgl)                           # (c, nn)
    for q in range(len(w)):
        JxW = w[q] * geom.detJ[sl]
        phi_u, G, div, phi_p, gpp = _basis_at(dim, geom, sl, lam[q])
        u_k, gu_k = _function_values(dim, phi_u, G, uc)
        u_o, gu_o = _function_values(dim, phi_u, G, uo)
        p_k = np.einsum("ck,k->c", uc, phi_p)
        gp_k = np.einsum("ck,ckd->cd", uc, gpp)
        lap_k = np.zeros((c, dim))
        for k in range(K):
            if comp[k] < dim:
                lap_k[:, comp[k]] += uc[:, k] * lapN[:, node[k]]
        f_new = np.zeros((c, dim)) if forcing is None else forcing[0][sl, q]
        f_old = np.zeros((c, dim)) if forcing is None else forcing[1][sl, q]
        f_th = th * f_new + (1.0 - th) * f_old
        conv_k = np.einsum("cde,ce->cd", gu_k, u_k)
        conv_o = np.einsum("cde,ce->cd", gu_o, u_o)
        # ---- residual (cpp:377-418)
        time_term = np.einsum("cd,kd->ck", u_k - u_o, phi_u) / dt
        conv_impl = th * np.einsum("cd,kd->ck", conv_k, phi_u)
        visc_impl = th * nu * np.einsum("cde,ckde->ck", gu_k, G)
        conv_expl = (1.0 - th) * np.einsum("cd,kd->ck", conv_o, phi_u)
        visc_expl = (1.0 - th) * nu * np.einsum("cde,ckde->ck", gu_o, G)
        pres_term = -p_k[:, None] * div
        div_term = -phi_p[None, :] * np.trace(gu_k, axis1=1, axis2=2)[:, None]
        b += (-time_term - conv_impl - visc_impl - conv_expl - visc_expl - pres_term - div_term) * JxW[:, None]
        b += np.einsum("cd,kd->ck", f_th, phi_u) * JxW[:, None]
        # ---- Jacobian (cpp:421-437)
        Gu = np.einsum("ckde,ce->ckd", G, u_k)                           # grad_phi_u[j] * u_k
        gukphi = np.einsum("cde,je->cjd", gu_k, phi_u)                   # grad u_k * phi_u[j]
        val = np.einsum("id,jd->ij", phi_u, phi_u)[None] / dt
        val = val + th * nu * np.einsum("cide,cjde->cij", G, G)
        val = val + th * (np.einsum("cjd,id->cij", Gu, phi_u) + np.einsum("cjd,id->cij", gukphi, phi_u))
        val = val - np.einsum("j,ci->cij", phi_p, div)
        val = val - np.einsum("i,cj->cij", phi_p, div)
        A += val * JxW[:, None, None]
        if p.use_supg:
            tau = _tau(p, np.sqrt((u_k ** 2).sum(1)), h)
            supg = tau[:, None, None] * Gu                               # cpp:452-453
            op = phi_u[None] / dt + Gu + gukphi                          # cpp:454-456
            A += np.einsum("cid,cjd->cij", supg, op + gpp) * JxW[:, None, None]
            A += p.gamma * np.einsum("ci,cj->cij", div, div) * JxW[:, None, None]
            strong = (u_k - u_o) / dt + conv_k + gp_k - nu * lap_k - f_th     # cpp:488-506
            b -= np.einsum("ckd,cd->ck", Gu, strong) * (tau * JxW)[:, None]   # cpp:508-509
        Mp += np.einsum("i,j->ij", phi_p, phi_p)[None] * JxW[:, None, None]
        Kp += np.einsum("cid,cjd->cij", gpp, gpp) * JxW[:, None, None]
    return A, b, Mp, Kp


# ---------------------------------------------------------------------------------
# AffineConstraints::distribute_local_to_global for pure Dirichlet lines
# ---------------------------------------------------------------------------------
def eliminate_local(cell_matrix, cell_rhs, cd, con):
    """Returns (matrix contributions (c,K,K) to add at (cd_i, cd_j), rhs contributions (c,K))."""
    isc = con.is_c[cd]                                  # (c, K)
    g = np.where(isc, con.val[cd], 0.0)
    K = cd.shape[1]
    out = cell_matrix.copy()
    rhs = None
    if cell_rhs is not None:
        rhs = cell_rhs - np.einsum("cij,cj->ci", cell_matrix, g)    # r_i - sum_c m_ic g_c
        rhs[isc] = 0.0
    out[isc[:, :, None] | isc[:, None, :]] = 0.0
    diag = np.abs(np.einsum("cii->ci", cell_matrix))
    avg = diag.sum(axis=1) / K
    new_diag = np.where(diag != 0.0, diag, avg[:, None])
    idx = np.arange(K)
    out[:, idx, idx] = np.where(isc, new_diag, out[:, idx, idx])
    return out, rhs


class Assembled:
    pass


def assemble(mesh, dm, pattern, p: Params, con, kind, vec_a, vec_b, with_pressure_matrices=True,
             forcing=None, chunk=2048):
    """Global assembly.  kind = 'linearized' (vec_a = solution_old, vec_b = solution_old_old)
    or 'newton' (vec_a = current_solution, vec_b = solution_old).
    Returns CSR values aligned with `pattern` = (rowptr, col) for A, Mp, Kp (full pattern,
    like the reference's three BlockSparseMatrix objects) and the rhs."""
    rowptr, col = pattern
    N = dm.n_dofs
    nnz = col.shape[0]
    geom = CellGeometry(mesh)
    A = np.zeros(nnz)
    Mp = np.zeros(nnz)
    Kp = np.zeros(nnz)
    b = np.zeros(N)
    C = mesh.n_cells
    allkeys = np.repeat(np.arange(N, dtype=np.int64), np.diff(rowptr)) * N + col   # sorted (CSR order)
    for s in range(0, C, chunk):
        sl = slice(s, min(C, s + chunk))
        if kind == "linearized":
            cm, cr, cmp_, ckp = cell_matrices_linearized(mesh, dm, p, vec_a, vec_b, sl, forcing, geom)
        else:
            cm, cr, cmp_, ckp = cell_matrices_newton(mesh, dm, p, vec_a, vec_b, sl, forcing, geom)
        cd = dm.cell_dofs[sl].astype(np.int64)
        K = cd.shape[1]
        m_out, r_out = eliminate_local(cm, cr, cd, con)
        ii = np.repeat(cd, K, axis=1).ravel()
        jj = np.tile(cd, (1, K)).ravel()
        want = ii * N + jj
        pos = np.searchsorted(allkeys, want)
        assert np.all(allkeys[pos] == want), "cell coupling missing from the sparsity pattern"
        np.add.at(A, pos, m_out.ravel())
        np.add.at(b, cd.ravel(), r_out.ravel())
        if with_pressure_matrices:
            mp_out, _ = eliminate_local(cmp_, None, cd, con)
            kp_out, _ = eliminate_local(ckp, None, cd, con)
            np.add.at(Mp, pos, mp_out.ravel())
            np.add.at(Kp, pos, kp_out.ravel())
    out = Assembled()
    out.A, out.b = A, b
    if with_pressure_matrices:
        Kp = Kp + 1e-6 * Mp          # pressure_stiffness.add(1e-6, pressure_mass)   cpp:536, 828
        out.Mp, out.Kp = Mp, Kp
    return out


def to_csr(pattern, values, N):
    rowptr, col = pattern
    return sp.csr_matrix((values, col, rowptr), shape=(N, N))


def pressure_blocks(mesh, dm, con):
    """(1,1) blocks of pressure_mass / pressure_stiffness (cpp:798-803, 812-829) as ONE compact CSR over the pressure
    DoFs: returns (ptr int64, col int32, Mp, Kp), identical to the (1,1) blocks of assemble(...).Mp / .Kp.
    M_p,ij = sum_q psi_i psi_j JxW, K_p,ij = sum_q grad psi_i . grad psi_j JxW, both through the matrix-only
    distribute_local_to_global (constrained rows / columns dropped, |m_cc| on the diagonal -- never zero here, so the
    average-diagonal branch of A.5 is not taken), then K_p += 1e-6 M_p.  Vectorised over the cells so that the
    multi-million-cell meshes of the CPU baseline do not need two more full-pattern value arrays."""
    dim = mesh.dim
    nv = dim + 1
    geom = CellGeometry(mesh)
    pts, w = fe.quadrature(dim)
    lam = fe.barycentric(pts)                                    # (Q, nv): P1 shape values
    mhat = np.einsum("q,qi,qj->ij", w, lam, lam)                 # reference mass matrix with the reference's rule
    node, comp = fe.local_dof_layout(dim)
    ploc = np.array([k for k in range(node.shape[0]) if comp[k] == dim])
    ploc = ploc[np.argsort(node[ploc])]                          # local pressure DoF of vertex 0..dim
    pid = dm.cell_dofs[:, ploc].astype(np.int64) - dm.n_u        # (C, nv)
    gl = geom.grad_lambda
    cm = geom.detJ[:, None, None] * mhat[None, :, :]
    ck = np.zeros_like(cm)
    for q in range(len(w)):                                      # same summation order as the cell loop
        ck += np.einsum("cid,cjd->cij", gl, gl) * (w[q] * geom.detJ)[:, None, None]
    isc = con.is_c[dm.n_u + pid]                                 # (C, nv)
    drop = isc[:, :, None] | isc[:, None, :]
    idx = np.arange(nv)
    out = []
    for loc in (cm, ck):
        o = np.where(drop, 0.0, loc)
        o[:, idx, idx] = np.where(isc, np.abs(loc[:, idx, idx]), o[:, idx, idx])
        out.append(o)
    ii = np.repeat(pid, nv, axis=1).ravel()
    jj = np.tile(pid, (1, nv)).ravel()
    n_p = dm.n_p
    Mp = sp.coo_matrix((out[0].ravel(), (ii, jj)), shape=(n_p, n_p)).tocsr()
    Kp = sp.coo_matrix((out[1].ravel(), (ii, jj)), shape=(n_p, n_p)).tocsr()
    Mp.sort_indices(); Kp.sort_indices()
    assert np.array_equal(Mp.indptr, Kp.indptr) and np.array_equal(Mp.indices, Kp.indices)
    return Mp.indptr.astype(np.int64), Mp.indices.astype(np.int32), Mp.data.copy(), Kp.data + 1e-6 * Mp.data
