"""Multi-GPU validation (one process per GPU, launch with torchrun): every rank's owned rows of the
assembled matrix / rhs against the oracle, the tight-tolerance solve against a sparse direct solve,
and the iteration count at the reference tolerance (partition-independent preconditioner).
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/gpu_multi.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import assemble as asm, dofs as odofs, postprocess as pp, solve as osolve  # noqa: E402
from tests.conftest import load_nsb, synthetic_state  # noqa: E402
from tools import meshgen  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nsb = load_nsb()
    holder = [nsb.Device.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(holder, src=0)
    mesh = meshgen.mesh_3d(lc_cyl=0.04, lc_global=0.15)
    dm = odofs.enumerate_dofs(mesh)
    pat = odofs.make_sparsity(dm)
    N = dm.n_dofs
    tc = pp.TEST_CASES["3D-2Z"]
    con = odofs.build_constraints(mesh, dm, pp.inlet_profile(3, tc["U_m"], False, 4.0, 1.0), pp.boundary_ids(3))
    un, unm1 = synthetic_state(dm, 3, tc["U_m"])
    p = asm.Params(dt=0.01, theta=0.5, nu=1e-3, use_supg=True)
    ref = asm.assemble(mesh, dm, pat, p, con, "linearized", un, unm1) if rank == 0 else None
    holder2 = [ref]
    dist.broadcast_object_list(holder2, src=0)
    ref = holder2[0]
    dev = nsb.Device(3, local)
    dev.comm_init(rank, world, holder[0])
    part = (np.arange(mesh.n_cells, dtype=np.int64) * world // mesh.n_cells).astype(np.int32)
    dev.upload_mesh(mesh.points, mesh.cells, dm.cell_dofs, dm.n_u, dm.n_p, part)
    nrows, nnz, nc = dev.sizes()
    dev.set_constraints(con.dofs, con.val[con.dofs])
    dev.set_params(0.01, 0.5, 1e-3, 1.0, 0.1, True, False)
    dev.set_vector(nsb.NSB_SOLUTION_OLD, un)
    dev.set_vector(nsb.NSB_SOLUTION_OLD_OLD, unm1)
    dev.assemble_linearized()
    rp, col = dev.pattern()
    gid = dev.row_gids()
    vals = dev.matrix_values()
    orp, ocol = pat
    errA = 0.0
    okpat = True
    for k in range(nrows):
        g = gid[k]
        okpat &= np.array_equal(col[rp[k]:rp[k + 1]].astype(np.int64), ocol[orp[g]:orp[g + 1]].astype(np.int64))
        errA = max(errA, np.abs(vals[rp[k]:rp[k + 1]] - ref.A[orp[g]:orp[g + 1]]).max())
    errA /= np.abs(ref.A).max()
    b = dev.get_vector(nsb.NSB_RHS)
    errb = np.abs(b - ref.b).max() / np.abs(ref.b).max()
    dev.assemble_pressure_matrices()
    ok, it, res = dev.solve(200, 1e-2, 150)
    ok2, it2, _ = dev.solve(3000, 1e-12, 150)
    xs = dev.get_vector(nsb.NSB_SOLUTION)
    out = f"[rank {rank}/{world}] rows {nrows} cells {nc} pattern_ok {okpat} A relerr {errA:.2e} b relerr {errb:.2e} | GMRES(1e-2) ok {ok} its {it} | tight ok {ok2} its {it2}"
    if rank == 0:
        A = asm.to_csr(pat, ref.A, N)
        xd = con.distribute(osolve.direct_solve(A, ref.b))
        out += f" field rel L2 vs direct {np.linalg.norm(xs - xd) / np.linalg.norm(xd):.2e}"
    print(out, flush=True)
    dist.barrier()
    dev.close()
    # ---- the C++ host class NavierStokes<3>(make_3D_2Z) on `world` GPUs: three time steps in parity mode
    from tools import msh
    path = "/tmp/gpu_multi_mesh_%d.bin" % rank
    msh.write_bin(path, mesh)
    holder3 = [nsb.Device.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(holder3, src=0)
    hs = nsb.HostSolver("3D-2Z", path, device=local, rank=rank, nranks=world, nccl_unique_id=holder3[0], gmres_tolerance=1e-12)
    hs.initialize()
    infos = [hs.step() for _ in range(3)]
    if rank == 0:
        o = osolve.Oracle(mesh, "3D-2Z", solver="direct")
        for k, info in enumerate(infos):
            ref_ = o.step()
            errs = {key: abs(info[key] - ref_[key]) / max(abs(ref_[key]), 1e-300) for key in ("cd", "cl", "dp")}
            print(f"[host class, {world} GPUs] step {k + 1}: gmres {info['gmres_iterations']} its, rel err vs oracle "
                  + ", ".join(f"{a}={b:.1e}" for a, b in errs.items()) + f"  (Cd={info['cd']:.6g} Cl={info['cl']:.3g} dP={info['dp']:.6g})", flush=True)
        x = hs.solution()
        print(f"[host class, {world} GPUs] field rel L2 vs oracle after 3 steps: {np.linalg.norm(x - o.current_solution) / np.linalg.norm(o.current_solution):.2e}", flush=True)
    dist.barrier()
    hs.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
