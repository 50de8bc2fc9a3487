"""Generates tests/golden/oracle_vectors.npz from the numpy oracle (python -m tools.make_golden_vectors).

These are REGRESSION pins of the oracle, not reference outputs: the reference cannot be built here
(no deal.II / Trilinos / MPI) and ships no golden data, so parity stays "unpinned" w.r.t. the
reference itself (see oracle/*.py headers and DESIGN.md)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import assemble as asm, dofs as odofs, postprocess as pp, solve as osolve  # noqa: E402
from tools import msh  # noqa: E402
from tests.conftest import synthetic_state  # noqa: E402


def main():
    out = {}
    m = msh.load_npz(os.path.join(ROOT, "tests/golden/mesh-2D.npz"))
    dm = odofs.enumerate_dofs(m)
    pat = odofs.make_sparsity(dm)
    ids = pp.boundary_ids(2)
    nu = pp.viscosity(2, 1.5, 100.0)
    con = odofs.build_constraints(m, dm, pp.inlet_profile(2, 1.5, False, 2.0, 1.0), ids)
    un, unm1 = synthetic_state(dm, 2, 1.5)
    p = asm.Params(dt=0.02, theta=0.5, nu=nu)
    a = asm.assemble(m, dm, pat, p, con, "linearized", un, unm1)
    sel = np.arange(0, pat[1].size, 997)
    out["lin_A_sel"], out["lin_A_sum"], out["lin_A_abs"] = a.A[sel], a.A.sum(), np.abs(a.A).sum()
    out["lin_b_sel"], out["lin_b_norm"] = a.b[::37], np.linalg.norm(a.b)
    out["Mp_sum"], out["Kp_abs"] = a.Mp.sum(), np.abs(a.Kp).sum()
    conN = odofs.build_constraints(m, dm, None, ids, homogeneous=True)
    pk = np.zeros(dm.n_dofs)
    pk[dm.n_u:] = 0.05 * np.random.default_rng(7).uniform(-1, 1, dm.n_p)
    pN = asm.Params(dt=0.1, theta=1.0, nu=pp.viscosity(2, 0.3, 20.0))
    n = asm.assemble(m, dm, pat, pN, conN, "newton", un + pk, unm1, with_pressure_matrices=False)
    out["newton_A_sel"], out["newton_A_abs"] = n.A[sel], np.abs(n.A).sum()
    out["newton_b_sel"], out["newton_b_norm"] = n.b[::37], np.linalg.norm(n.b)
    out["sel_stride"] = np.array([997, 37])
    # trajectories with direct solves (parity mode)
    o = osolve.Oracle(m, "2D-2", solver="direct")
    traj = [o.step() for _ in range(4)]
    out["traj_2D2"] = np.array([[t["time"], t["cd"], t["cl"], t["dp"]] for t in traj])
    o = osolve.Oracle(m, "2D-1", solver="direct")
    t = o.step()
    out["traj_2D1"] = np.array([[t["time"], t["cd"], t["cl"], t["dp"], t["newton_iters"]]])
    np.savez_compressed(os.path.join(ROOT, "tests/golden/oracle_vectors.npz"), **out)
    for k, v in out.items():
        print(k, np.asarray(v).shape)


if __name__ == "__main__":
    main()
