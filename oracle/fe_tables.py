"""ORACLE (test infrastructure, not product code) -- finite-element tables.

PARITY UNPINNED: the reference (gdonninelli/Navier-Stokes_equations) delegates all of
this to deal.II (un-vendored, >= 9.3.1, `common/cmake-common.cmake:28`), which cannot be
built here, and ships no tests / golden vectors.  What follows restates deal.II's
*documented* conventions; the pins are analytic identities
(tests/test_oracle_pins.py).

  FESystem(FE_SimplexP(2)^dim, FE_SimplexP(1))    reference NavierStokes.hpp:429-432
  QGaussSimplex<dim>(3)                           reference NavierStokes.hpp:433
  QGaussSimplex<dim-1>(3) on faces                reference NavierStokes.cpp:924
  MappingFE(FE_SimplexP(1)) = affine map          reference NavierStokes.hpp:435
"""
import numpy as np

# ---------------------------------------------------------------------------------
# reference simplices (deal.II ReferenceCells::Triangle / Tetrahedron)
#   triangle   vertices (0,0),(1,0),(0,1);          lines (0,1),(1,2),(2,0)
#   tetrahedron vertices (0,0,0),(1,0,0),(0,1,0),(0,0,1); lines (0,1),(1,2),(2,0),(0,3),(1,3),(2,3)
#   faces: triangle face f = line f; tetrahedron faces (0,1,2),(1,0,3),(0,2,3),(2,1,3)
# ---------------------------------------------------------------------------------
LINES = {
    2: np.array([[0, 1], [1, 2], [2, 0]], dtype=np.int32),
    3: np.array([[0, 1], [1, 2], [2, 0], [0, 3], [1, 3], [2, 3]], dtype=np.int32),
}
FACES = {
    2: np.array([[0, 1], [1, 2], [2, 0]], dtype=np.int32),
    3: np.array([[0, 1, 2], [1, 0, 3], [0, 2, 3], [2, 1, 3]], dtype=np.int32),
}


def n_nodes(dim):
    """P2 nodes per cell: vertices + lines."""
    return (dim + 1) + LINES[dim].shape[0]


def dofs_per_cell(dim):
    return dim * n_nodes(dim) + (dim + 1)


def local_dof_layout(dim):
    """FESystem local ordering: per vertex [u_0..u_{d-1}, p], then per line [u_0..u_{d-1}].
    Returns (node[k], comp[k]) with comp == dim for pressure; node indexes the P2 node
    (0..dim vertices, then lines)."""
    nv = dim + 1
    node, comp = [], []
    for v in range(nv):
        for c in range(dim + 1):
            node.append(v)
            comp.append(c)
    for l in range(LINES[dim].shape[0]):
        for c in range(dim):
            node.append(nv + l)
            comp.append(c)
    return np.array(node, np.int32), np.array(comp, np.int32)


# ---------------------------------------------------------------------------------
# quadrature.  deal.II 9.3/9.4 `QGaussSimplex<dim>(3)` hard-coded tables
# (SURVEY.md A.3): 2-D 7-point Radon rule with ~13-digit constants (centroid literally
# 0.3333333333330); 3-D 10-point degree-3 Keast rule.  Points are reference coordinates
# (x,y[,z]); weights sum to 1/2 resp. 1/6.
# ---------------------------------------------------------------------------------
def quadrature(dim):
    if dim == 2:
        pts = np.array([
            [0.3333333333330, 0.3333333333330],
            [0.7974269853530, 0.1012865073230],
            [0.1012865073230, 0.7974269853530],
            [0.1012865073230, 0.1012865073230],
            [0.0597158717898, 0.4701420641050],
            [0.4701420641050, 0.0597158717898],
            [0.4701420641050, 0.4701420641050]])
        w = 0.5 * np.array([0.225, 0.125939180545, 0.125939180545, 0.125939180545,
                            0.132394152789, 0.132394152789, 0.132394152789])
        return pts, w
    if dim == 3:
        a, b = 0.5684305841968444, 0.1438564719343852
        pts = np.array([
            [a, b, b], [b, b, b], [b, b, a], [b, a, b],
            [0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0],
            [0.5, 0.0, 0.0], [0.0, 0.5, 0.0], [0.0, 0.0, 0.5]])
        w = np.array([0.2177650698804054] * 4 + [0.0214899534130631] * 6) / 6.0
        return pts, w
    raise ValueError(dim)


def face_quadrature(dim):
    """QGaussSimplex<dim-1>(3): 3-point Gauss-Legendre on [0,1] (2-D faces), the 7-point
    triangle rule above (3-D faces)."""
    if dim == 2:
        x, w = np.polynomial.legendre.leggauss(3)
        return (0.5 * (x + 1.0)).reshape(-1, 1), 0.5 * w
    return quadrature(2)


# ---------------------------------------------------------------------------------
# shape functions in barycentric form.  lambda_0 = 1 - sum(x), lambda_k = x_{k-1}.
# P2: vertex i: lambda_i (2 lambda_i - 1); line (i,j): 4 lambda_i lambda_j.  P1: lambda_i.
# ---------------------------------------------------------------------------------
def barycentric(pts):
    pts = np.atleast_2d(pts)
    return np.concatenate([1.0 - pts.sum(axis=1, keepdims=True), pts], axis=1)


def p2_values(dim, lam):
    """lam (Q, dim+1) -> N (Q, n_nodes), dN/dlambda (Q, n_nodes, dim+1)."""
    nv = dim + 1
    Q = lam.shape[0]
    nn = n_nodes(dim)
    N = np.zeros((Q, nn))
    dN = np.zeros((Q, nn, nv))
    for i in range(nv):
        N[:, i] = lam[:, i] * (2.0 * lam[:, i] - 1.0)
        dN[:, i, i] = 4.0 * lam[:, i] - 1.0
    for l, (i, j) in enumerate(LINES[dim]):
        N[:, nv + l] = 4.0 * lam[:, i] * lam[:, j]
        dN[:, nv + l, i] = 4.0 * lam[:, j]
        dN[:, nv + l, j] = 4.0 * lam[:, i]
    return N, dN


def p2_second(dim):
    """d2N/dlambda_k dlambda_l, constant: (n_nodes, dim+1, dim+1)."""
    nv = dim + 1
    H = np.zeros((n_nodes(dim), nv, nv))
    for i in range(nv):
        H[i, i, i] = 4.0
    for l, (i, j) in enumerate(LINES[dim]):
        H[nv + l, i, j] = 4.0
        H[nv + l, j, i] = 4.0
    return H


def support_points_ref(dim):
    """Reference support points of the P2 nodes (vertices, then line midpoints)."""
    nv = dim + 1
    V = np.zeros((nv, dim))
    for k in range(1, nv):
        V[k, k - 1] = 1.0
    mids = 0.5 * (V[LINES[dim][:, 0]] + V[LINES[dim][:, 1]])
    return np.concatenate([V, mids], axis=0)
