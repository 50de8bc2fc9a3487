"""CPU checks, against the oracle's cell matrices, of the algebra the assembly and the two-level velocity cycle rest on
(csrc/assemble.cuh, csrc/twolevel.cuh): (1) every velocity-velocity cell block of the linearised system is
delta_cd S_ab + gamma G^{cd}_ab with G = |J| g^T Khat g, for every u* regime and with SUPG; (2) the closed-form grad-div
action (div x_h is P1 on the cell) equals G x.  The GPU tests test_linearized_assembly_parity and
test_coarse_operator_is_the_galerkin_product check the kernels themselves."""
import math

import numpy as np
import pytest

from oracle import assemble as asm, dofs as odofs, fe_tables as fe
from tests.conftest import synthetic_state


def _ldof(dim, a, c):
    nv = dim + 1
    return a * (dim + 1) + c if a < nv else nv * (dim + 1) + (a - nv) * dim + c


def _khat_full(dim):
    pts, w = fe.quadrature(dim)
    N, dN = fe.p2_values(dim, fe.barycentric(pts))
    return np.einsum("q,qak,qbl->abkl", w, dN, dN)


@pytest.mark.parametrize("which,supg,first", [("2d", False, False), ("2d", True, False), ("3d", True, False), ("3d", True, True)])
def test_velocity_cell_blocks_are_scalar_plus_graddiv(golden_mesh, small_3d_mesh, which, supg, first):
    mesh = golden_mesh("mesh-2D") if which == "2d" else small_3d_mesh
    dim = mesh.dim
    dm = odofs.enumerate_dofs(mesh)
    un, unm1 = synthetic_state(dm, dim, 1.5)
    p = asm.Params(dt=0.01, theta=0.5, nu=1e-3, use_supg=supg, first_step=first)
    sl = slice(0, 40)
    cm, _, _, _ = asm.cell_matrices_linearized(mesh, dm, p, un, unm1, sl)
    geom = asm.CellGeometry(mesh)
    gl, detJ = geom.grad_lambda[sl], np.abs(geom.detJ[sl])
    G = np.einsum("e,ekc,abkl,eld->eabcd", detJ, gl, _khat_full(dim), gl)
    gamma = p.gamma if supg else 0.0
    nn = fe.n_nodes(dim)
    worst_off = worst_spread = 0.0
    for e in range(40):
        for a in range(nn):
            for b in range(nn):
                blk = np.array([[cm[e, _ldof(dim, a, c), _ldof(dim, b, d)] for d in range(dim)] for c in range(dim)])
                R = blk - gamma * G[e, a, b]
                scale = np.abs(cm[e]).max()
                worst_off = max(worst_off, np.abs(R - np.diag(np.diag(R))).max() / scale)
                worst_spread = max(worst_spread, np.abs(np.diag(R) - np.diag(R).mean()).max() / scale)
    assert worst_off < 1e-13 and worst_spread < 1e-13


@pytest.mark.parametrize("dim", [2, 3])
def test_closed_form_graddiv_action(dim):
    """w_a^s = sum_b sum_t Khat[a][b][s][t] z_b^t  from the four P1 coefficients of div x_h (ebe.cuh)."""
    nv, nn = dim + 1, fe.n_nodes(dim)
    idx = [(v, v) for v in range(nv)] + [tuple(l) for l in fe.LINES[dim]]
    K4 = _khat_full(dim)
    rng = np.random.default_rng(0)
    z = rng.standard_normal((nn, 2))
    z[:nv, 1] = 0.0
    w_ref = np.zeros((nn, 2))
    for a in range(nn):
        for s in range(2 if a >= nv else 1):
            for b in range(nn):
                for t in range(2 if b >= nv else 1):
                    w_ref[a, s] += K4[a, b, idx[a][s], idx[b][t]] * z[b, t]
    ref = 1.0 / math.factorial(dim)
    m2, m1 = ref / ((dim + 1) * (dim + 2)), ref / (dim + 1)
    D, Z = np.zeros(nv), z[:nv, 0].sum()
    for b in range(nn):
        i, j = idx[b]
        if b < nv:
            D[i] += z[b, 0]
        else:
            D[j] += z[b, 0]
            D[i] += z[b, 1]
    D = 4.0 * D - Z
    SD = D.sum()
    M = m2 * (SD + D)
    w = np.zeros((nn, 2))
    for a in range(nn):
        i, j = idx[a]
        if a < nv:
            w[a, 0] = 4.0 * M[i] - m1 * SD
        else:
            w[a, 0], w[a, 1] = 4.0 * M[j], 4.0 * M[i]
    # the 2-D rule carries 13-digit constants, hence 1e-11
    assert np.abs(w - w_ref).max() / np.abs(w_ref).max() < 1e-11
