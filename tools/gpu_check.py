"""Developer harness: CUDA path vs the numpy oracle on one GPU (prints errors, no asserts).
Usage: python -m tools.gpu_check [2d|3d|all]"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import assemble as asm, dofs as odofs, postprocess as pp, solve as osolve  # noqa: E402
from tools import meshgen, msh  # noqa: E402


def load_nsb():
    import nsb200
    return nsb200


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def check(mesh, case, tag):
    nsb = load_nsb()
    tc = pp.TEST_CASES[case]
    dim = mesh.dim
    t0 = time.time()
    dm = odofs.enumerate_dofs(mesh)
    pat = odofs.make_sparsity(dm)
    print(f"[{tag}] cells {mesh.n_cells} dofs {dm.n_u}+{dm.n_p} nnz {pat[1].size}  oracle setup {time.time()-t0:.1f}s")
    N = dm.n_dofs
    dev = nsb.Device(dim)
    t0 = time.time()
    dev.upload_mesh(mesh.points, mesh.cells, dm.cell_dofs, dm.n_u, dm.n_p)
    print(f"[{tag}] upload_mesh {time.time()-t0:.2f}s sizes {dev.sizes()}")
    rp, col = dev.pattern()
    print(f"[{tag}] pattern bit-exact: rowptr {np.array_equal(rp, pat[0])} col {np.array_equal(col.astype(np.int64), pat[1].astype(np.int64))}")
    ids = pp.boundary_ids(dim)
    nu = pp.viscosity(dim, tc["U_m"], tc["Re"])
    t_now = 1.0
    inlet = pp.inlet_profile(dim, tc["U_m"], tc["time_dep"], tc["T_ramp"], t_now)
    con = odofs.build_constraints(mesh, dm, inlet, ids)
    cd = con.dofs
    dev.set_constraints(cd, con.val[cd])
    # synthetic state
    full = pp.inlet_profile(dim, tc["U_m"], False, 0.0, 0.0)
    base = np.zeros(N)
    base[:dm.n_u] = full(dm.support_points[:dm.n_u], dm.component[:dm.n_u])
    un = base * (1 + 0.1 * np.random.default_rng(1234).uniform(-1, 1, N))
    unm1 = base * (1 + 0.1 * np.random.default_rng(1235).uniform(-1, 1, N))
    pk = np.zeros(N)
    pk[dm.n_u:] = 0.05 * np.random.default_rng(7).uniform(-1, 1, dm.n_p)
    dt = 0.01 if dim == 3 else 0.02
    for (theta, fo) in ((0.5, False), (1.0, True)):
        p = asm.Params(dt=dt, theta=theta, nu=nu, use_supg=tc["supg"], first_step=fo)
        dev.set_params(dt, theta, nu, 1.0, 0.1, tc["supg"], fo)
        dev.set_vector(nsb.NSB_SOLUTION_OLD, un)
        dev.set_vector(nsb.NSB_SOLUTION_OLD_OLD, unm1)
        t0 = time.time()
        out = asm.assemble(mesh, dm, pat, p, con, "linearized", un, unm1, with_pressure_matrices=(theta == 0.5))
        t_or = time.time() - t0
        dev.assemble_linearized()
        dev.synchronize()
        A = dev.matrix_values()
        b = dev.get_vector(nsb.NSB_RHS)
        print(f"[{tag}] linearized theta={theta} first_order={fo}: A relerr {relerr(A, out.A):.2e}  b relerr {relerr(b, out.b):.2e}  (oracle {t_or:.1f}s)")
        if theta == 0.5:
            ref = out
    # Newton
    ucur = un + pk
    pN = asm.Params(dt=dt, theta=1.0, nu=nu, use_supg=tc["supg"])
    conN = odofs.build_constraints(mesh, dm, None, ids, homogeneous=True)
    dev.set_constraints(conN.dofs, conN.val[conN.dofs])
    dev.set_params(dt, 1.0, nu, 1.0, 0.1, tc["supg"], False)
    dev.set_vector(nsb.NSB_CURRENT_SOLUTION, ucur)
    dev.set_vector(nsb.NSB_SOLUTION_OLD, unm1)
    outN = asm.assemble(mesh, dm, pat, pN, conN, "newton", ucur, unm1, with_pressure_matrices=False)
    dev.assemble_newton()
    A = dev.matrix_values()
    b = dev.get_vector(nsb.NSB_RHS)
    print(f"[{tag}] newton: A relerr {relerr(A, outN.A):.2e}  b relerr {relerr(b, outN.b):.2e}")
    # back to the linearised system for SpMV / solve
    dev.set_constraints(cd, con.val[cd])
    dev.set_params(dt, 0.5, nu, 1.0, 0.1, tc["supg"], False)
    dev.set_vector(nsb.NSB_SOLUTION_OLD, un)
    dev.set_vector(nsb.NSB_SOLUTION_OLD_OLD, unm1)
    dev.assemble_linearized()
    t0 = time.time()
    dev.assemble_pressure_matrices()
    print(f"[{tag}] pressure matrices + AMG setup {time.time()-t0:.2f}s")
    n_u = dm.n_u
    Aor = asm.to_csr(pat, ref.A, N)
    for which, name, full_ in ((0, "Mp", ref.Mp), (1, "Kp", ref.Kp)):
        prp, pcol, pval = dev.pressure_matrix(which)
        Mdev = sp.csr_matrix((pval, pcol, prp), shape=(dm.n_p, dm.n_p))
        Mor = asm.to_csr(pat, full_, N)[n_u:, n_u:]
        print(f"[{tag}] {name} relerr {abs(Mdev - Mor).max() / abs(Mor).max():.2e}")
    x = np.random.default_rng(42).uniform(-1, 1, N)
    y = dev.spmv(x)
    print(f"[{tag}] spmv relerr {relerr(y, Aor @ x):.2e}")
    # reference-tolerance solve and tight solve
    ok, it, res = dev.solve(200, 1e-2, 150)
    print(f"[{tag}] solve tol 1e-2: ok {ok} iters {it} res {res:.3e}")
    t0 = time.time()
    ok, it, res = dev.solve(2000, 1e-12, 150)
    dev.synchronize()
    xs = dev.get_vector(nsb.NSB_SOLUTION)
    print(f"[{tag}] solve tol 1e-12: ok {ok} iters {it} res {res:.3e} {time.time()-t0:.2f}s")
    if N < 200000:
        xd = con.distribute(osolve.direct_solve(Aor, ref.b))
        print(f"[{tag}] field parity vs direct solve: rel L2 {np.linalg.norm(xs - xd) / np.linalg.norm(xd):.2e}")
    r = ref.b - Aor @ np.where(con.is_c, 0.0, xs)
    print(f"[{tag}] true residual rel {np.linalg.norm(r) / np.linalg.norm(ref.b):.2e}")
    print(f"[{tag}] launches {dev.launch_count()}")
    dev.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("2d", "all"):
        check(msh.load_npz(os.path.join(ROOT, "tests/golden/mesh-2D.npz")), "2D-2", "mesh-2D")
    if what in ("3d", "all"):
        m = meshgen.mesh_3d(lc_cyl=0.04, lc_global=0.15)
        check(m, "3D-2Z", "3d-small")
