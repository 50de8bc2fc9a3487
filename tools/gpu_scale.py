"""Developer harness: first-order performance numbers of the CUDA path at scale.
Usage: python -m tools.gpu_scale LEVEL [steps]      (LEVEL in 5,10,20)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dofs as odofs, postprocess as pp  # noqa: E402
from tools import meshgen  # noqa: E402
from tools.gpu_check import load_nsb  # noqa: E402


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    nsb = load_nsb()
    import bench
    t0 = time.time()
    mesh_file = bench.get_mesh_file(level)
    hs = nsb.HostSetup(mesh_file, 3)
    pts, cells = hs.mesh()
    cell_dofs = hs.cell_dofs()
    sp_pts, comp = hs.support_points()
    cd, cvals = hs.constraints("3D-2Z", 1.0)
    N = hs.n_dofs
    print(f"mesh-3D-{level}: cells {hs.n_cells} dofs {hs.n_u}+{hs.n_p}  host setup {time.time()-t0:.1f}s  lib {os.path.basename(nsb.LIB_PATH)}", flush=True)
    tc = pp.TEST_CASES["3D-2Z"]
    dev = nsb.Device(3)
    t0 = time.time()
    dev.upload_mesh(pts, cells, cell_dofs, hs.n_u, hs.n_p)
    nrows, nnz, nc = dev.sizes()
    print(f"upload_mesh {time.time()-t0:.1f}s  rows {nrows} nnz {nnz} ({nnz/nrows:.1f}/row)", flush=True)
    nu = pp.viscosity(3, tc["U_m"], tc["Re"])
    dev.set_constraints(cd, cvals)
    un, unm1 = bench.synthetic_state(sp_pts, comp, hs.n_u)
    dev.set_params(0.01, 0.5, nu, 1.0, 0.1, True, False)
    dev.set_vector(nsb.NSB_SOLUTION_OLD, un)
    dev.set_vector(nsb.NSB_SOLUTION_OLD_OLD, unm1)
    dev.assemble_linearized()
    dev.synchronize()
    t0 = time.time()
    dev.assemble_pressure_matrices()
    print(f"pressure matrices + AMG {time.time()-t0:.1f}s", flush=True)
    asm_bytes = 8 * nnz + 8 * nrows + 8 * 3 * 4 * nc + 20 * 34 * nc
    spmv_bytes = 12 * nnz + 16 * nrows + 4 * (nrows + 1)
    sweeps = [dict()]
    if len(sys.argv) > 3:
        sweeps = [dict(poly_target=t) for t in (0.05, 0.08, 0.12)]
    for deg in sweeps:
        dev.set_solver_opts(**deg)
        dev.profile_enable(True)
        dev.profile_reset()
        for s in range(steps):
            dev.timer_start()
            dev.assemble_linearized()
            ta = dev.timer_stop()
            dev.timer_start()
            ok, it, res = dev.solve(200, 1e-2, 150)
            ts = dev.timer_stop()
            print(f"target {deg} step {s}: assemble {ta:.2f} ms  solve {ts:.1f} ms  iters {it} ok {ok} res {res:.2e} {dev.solver_info()}", flush=True)
        prof = dev.profile()
        for k, (ms, n) in prof.items():
            if n:
                extra = ""
                if k == "spmv":
                    extra = f"  -> {spmv_bytes / (ms / n * 1e-3) / 1e9:.0f} GB/s algorithmic"
                if k == "asm_rows":
                    extra = f"  -> {asm_bytes / (ms / n * 1e-3) / 1e9:.0f} GB/s algorithmic (whole assembly bytes)"
                print(f"   {k:12s} {ms:10.2f} ms / {n:6d} launches = {ms/n:8.3f} ms{extra}")
        dev.profile_enable(False)
    dev.close()


if __name__ == "__main__":
    main()
