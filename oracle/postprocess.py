"""ORACLE (test infrastructure, not product code) -- drag / lift / pressure difference and
the inlet profile / test-case table.

PARITY UNPINNED by the reference (no tests or stored outputs).  Restates
  compute_lift_drag              reference src/classes/NavierStokes.cpp:913-1011
  compute_pressure_difference    reference src/classes/NavierStokes.cpp:871-912
  BenchmarkInletVelocity::value  reference src/classes/TestCases.hpp:29-75
  TestCases::make_*              reference src/classes/TestCases.hpp:101-306
"""
import numpy as np
from . import fe_tables as fe

D_CYL = 0.1     # NavierStokes.hpp:494
H_CH = 0.41     # NavierStokes.hpp:497


def inlet_profile(dim, U_m, time_dependent, T_ramp, t):
    """Returns f(points, comp) -> values, BenchmarkInletVelocity at time t (TestCases.hpp:29-75)."""
    H = H_CH

    def fn(pts, comp):
        pts = np.atleast_2d(pts)
        if dim == 2:
            y = pts[:, 1]
            prof = 4.0 * U_m * y * (H - y) / (H * H)
            flow = 0
        else:
            x, y = pts[:, 0], pts[:, 1]
            prof = 16.0 * U_m * x * y * (H - x) * (H - y) / (H * H * H * H)
            flow = 2
        if time_dependent:
            prof = prof * np.sin(np.pi * t / 8.0)
        if T_ramp > 0.0 and t < T_ramp:
            prof = prof * (0.5 * (1.0 - np.cos(np.pi * t / T_ramp)))
        return np.where(np.asarray(comp) == flow, prof, 0.0)

    return fn


def default_deltat(Re):
    # NavierStokes.hpp:368-375
    if Re <= 20:
        return 0.1
    if Re <= 50:
        return 0.05
    if Re <= 100:
        return 0.02
    if Re <= 150:
        return 0.01
    return 0.005


# name -> dict; mirrors TestCases.hpp:101-306
TEST_CASES = {
    "2D-1": dict(dim=2, Re=20.0, U_m=0.3, T=10.0, deltat=-1.0, scheme="BE", method="newton", time_dep=False, T_ramp=1.0, supg=False),
    "2D-2": dict(dim=2, Re=100.0, U_m=1.5, T=8.0, deltat=-1.0, scheme="CN", method="linearized", time_dep=False, T_ramp=2.0, supg=False),
    "2D-3": dict(dim=2, Re=100.0, U_m=1.5, T=8.0, deltat=-1.0, scheme="CN", method="linearized", time_dep=True, T_ramp=0.0, supg=False),
    "3D-1Z": dict(dim=3, Re=20.0, U_m=0.45, T=10.0, deltat=-1.0, scheme="BE", method="newton", time_dep=False, T_ramp=0.0, supg=True),
    "3D-2Z": dict(dim=3, Re=100.0, U_m=2.25, T=8.0, deltat=0.01, scheme="CN", method="linearized", time_dep=False, T_ramp=4.0, supg=True),
    "3D-3Z": dict(dim=3, Re=100.0, U_m=2.25, T=8.0, deltat=0.01, scheme="CN", method="linearized", time_dep=True, T_ramp=0.0, supg=True),
}


def boundary_ids(dim):
    # NavierStokes.hpp:518-521
    return dict(inlet=101, outlet=102, wall=103 if dim == 2 else 104, cylinder=104 if dim == 2 else 103)


def viscosity(dim, U_m, Re):
    # NavierStokes.cpp:64-70
    U_mean = (2.0 / 3.0) * U_m if dim == 2 else (4.0 / 9.0) * U_m
    return U_mean * D_CYL / Re


def _face_table(mesh):
    """boundary faces with a given tag -> (cell, local face) pairs, in cell order."""
    dim = mesh.dim
    V = mesh.n_vertices
    faces = fe.FACES[dim]
    cells = mesh.cells.astype(np.int64)

    def key(fv):
        s = np.sort(fv, axis=-1)
        k = s[..., 0]
        for a in range(1, dim):
            k = k * V + s[..., a]
        return k

    ck = key(cells[:, faces])                                     # (C, nf)
    return ck, key


def lift_drag(mesh, dm, solution, nu, rho, U_m, cylinder_id):
    """C_D, C_L by face quadrature of -(sigma n) over cylinder faces (cpp:913-1011)."""
    dim = mesh.dim
    ck, key = _face_table(mesh)
    bk = key(mesh.faces[mesh.face_tag == cylinder_id].astype(np.int64))
    hit = np.isin(ck, bk)
    cell_idx, face_idx = np.nonzero(hit)                           # cell order, then face order
    fq, fw = fe.face_quadrature(dim)
    node, comp = fe.local_dof_layout(dim)
    faces = fe.FACES[dim]
    nv = dim + 1
    Vref = np.zeros((nv, dim))
    for k in range(1, nv):
        Vref[k, k - 1] = 1.0
    force = np.zeros(dim)
    X = mesh.points[mesh.cells]
    for c, f in zip(cell_idx, face_idx):
        Xc = X[c]
        J = (Xc[1:] - Xc[:1]).T
        Jinv = np.linalg.inv(J)
        gl = np.vstack([-Jinv.sum(axis=0), Jinv])
        fv = faces[f]
        # reference points on the face: vertex0 + sum_k xi_k (vertex_k - vertex0)
        ref = Vref[fv[0]][None, :] + fq @ (Vref[fv[1:]] - Vref[fv[0]][None, :])
        lam = fe.barycentric(ref)
        N, dN = fe.p2_values(dim, lam)
        gradN = np.einsum("qnk,kd->qnd", dN, gl)
        # outward normal of the cell and surface measure
        P = Xc[fv]
        if dim == 2:
            t = P[1] - P[0]
            area = np.linalg.norm(t)
            n = np.array([t[1], -t[0]]) / area
            meas = area                       # reference face [0,1]: weights sum to 1
        else:
            cr = np.cross(P[1] - P[0], P[2] - P[0])
            area2 = np.linalg.norm(cr)
            n = cr / area2
            meas = area2                      # reference triangle area 1/2, weights sum to 1/2
        opp = [v for v in range(nv) if v not in fv][0]
        if np.dot(n, Xc[opp] - P[0]) > 0:
            n = -n
        loc = solution[dm.cell_dofs[c]]
        for q in range(len(fw)):
            JxW = fw[q] * meas
            grad_u = np.zeros((dim, dim))
            p = 0.0
            for k in range(len(node)):
                if comp[k] < dim:
                    grad_u[comp[k], :] += loc[k] * gradN[q, node[k], :]
                else:
                    p += loc[k] * lam[q, node[k]]
            stress = -p * np.eye(dim) + rho * nu * (grad_u + grad_u.T)
            force += -(stress @ n) * JxW
    U_mean = (2.0 / 3.0) * U_m if dim == 2 else (4.0 / 9.0) * U_m
    ref_area = D_CYL if dim == 2 else D_CYL * H_CH
    den = 0.5 * rho * U_mean * U_mean * ref_area
    if dim == 2:
        return force[0] / den, force[1] / den
    return force[2] / den, force[1] / den


def point_value_pressure(mesh, dm, solution, pt, tol=1e-10):
    """VectorTools::point_value for the pressure component; None if no cell contains pt."""
    dim = mesh.dim
    X = mesh.points[mesh.cells]
    J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))
    ref = np.linalg.solve(J, (np.asarray(pt)[None, :] - X[:, 0, :])[:, :, None])[:, :, 0]
    lam = np.concatenate([1.0 - ref.sum(1, keepdims=True), ref], axis=1)
    inside = np.all(lam >= -tol, axis=1)
    idx = np.nonzero(inside)[0]
    if idx.size == 0:
        return None
    c = idx[0]
    pd = dm.cell_dofs[c][[v * (dim + 1) + dim for v in range(dim + 1)]]
    return float(lam[c] @ solution[pd])


def pressure_difference(mesh, dm, solution):
    # cpp:871-912
    if mesh.dim == 2:
        a, b = (0.15, 0.2), (0.25, 0.2)
    else:
        a, b = (0.205, 0.2, 0.40), (0.205, 0.2, 0.50)
    pa = point_value_pressure(mesh, dm, solution, a)
    pb = point_value_pressure(mesh, dm, solution, b)
    return (pa if pa is not None else 0.0) - (pb if pb is not None else 0.0)
