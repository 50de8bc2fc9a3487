"""Developer harness: solver-option sweep on one GPU at a given mesh level (prints GMRES iterations, operator
applications and milliseconds per solve for every option set).  python tools/gpu_tune.py --level 20 'velocity_cycle=2,smoother_degree=6' ..."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nsb200 as nsb  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=10)
    ap.add_argument("sets", nargs="*")
    a = ap.parse_args()
    hs = nsb.HostSetup(bench.get_mesh_file(a.level), 3)
    pts, cells = hs.mesh()
    sp_pts, comp = hs.support_points()
    cdofs, cvals = hs.constraints("3D-2Z", 1.0)
    dev = nsb.Device(3, 0)
    dev.upload_mesh(pts, cells, hs.cell_dofs(), hs.n_u, hs.n_p)
    un, unm1 = bench.synthetic_state(sp_pts, comp, hs.n_u)
    dev.set_constraints(cdofs, cvals)
    dev.set_params(0.01, 0.5, 1e-3, 1.0, 0.1, True, False)
    dev.set_vector(nsb.NSB_SOLUTION_OLD, un)
    dev.set_vector(nsb.NSB_SOLUTION_OLD_OLD, unm1)
    dev.assemble_linearized()
    dev.assemble_pressure_matrices()
    for spec in a.sets or ["velocity_cycle=2"]:
        kw = {}
        for item in spec.split(","):
            k, v = item.split("=")
            kw[k] = float(v) if "." in v else int(v)
        dev.set_solver_opts(**kw)
        dev.assemble_linearized()
        dev.solve(200, 1e-2, 150)
        dev.profile_enable(True)
        dev.profile_reset()
        dev.synchronize()
        t0 = time.time()
        ok, it, res = dev.solve(200, 1e-2, 150)
        dev.synchronize()
        ms = (time.time() - t0) * 1e3
        pr = dev.profile()
        dev.profile_enable(False)
        print("%-70s ok %s its %3d  %7.1f ms  vel apps %4d (%.0f ms)  coarse %.0f ms  orth %.0f ms  spmv %.0f ms | %s"
              % (spec, ok, it, ms, pr["spmv_vel"][1], pr["spmv_vel"][0], pr["coarse"][0], pr["orth"][0], pr["spmv"][0], dev.velocity_pc_info()), flush=True)
    dev.close()


if __name__ == "__main__":
    main()
