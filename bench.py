#!/usr/bin/env python
"""bench.py -- time-steps/s of the B200 hot path (assembly of the linearised Taylor-Hood system +
block-preconditioned GMRES) on the mesh-3D-20-equivalent, with the roofline of the dominant kernel,
an end-to-end number through the C ABI with host buffers, and a CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--level 20] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path -- nsb_assemble_linearized +
nsb_solve with the reference's stopping rule (GMRES(150), 1e-2 * ||b||, <= 200 iterations) -- on the
synthetic state of SURVEY.md 8(d) (3D-2Z parameters: dt 0.01, theta 0.5, nu 1e-3, SUPG + grad-div).
`value` keeps all inputs resident in HBM; `e2e` pushes u^n, u^{n-1} and the constraints from host buffers
and reads the solution back every step.  Inputs are far larger than L2 (12.8 GB of matrix values),
so no explicit L2 flush is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
CACHE = os.path.join(ROOT, "meshes_cache")
CASE = "3D-2Z"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def get_mesh_file(level):
    """mesh-3D-<level>-equivalent as a binary dump for the C++ host reader (generated once, cached)."""
    from tools import meshgen, msh
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, "mesh-3D-%d.bin" % level)
    if not os.path.exists(path):
        t0 = time.time()
        m = meshgen.mesh_3d(level)
        msh.write_bin(path + ".tmp", m)
        os.replace(path + ".tmp", path)
        log("[bench] generated mesh-3D-%d-equivalent: %d cells, %d vertices in %.1f s" % (level, m.n_cells, m.n_vertices, time.time() - t0))
    return path


def synthetic_state(pts, comp, n_u, U_m=2.25, H=0.41):
    """u^n = inlet paraboloid * (1 + 0.1 xi), u^{n-1} likewise (seeds 1234 / 1235), p = 0."""
    N = pts.shape[0]
    prof = 16.0 * U_m * pts[:, 0] * pts[:, 1] * (H - pts[:, 0]) * (H - pts[:, 1]) / H ** 4
    base = np.where((comp == 2) & (np.arange(N) < n_u), prof, 0.0)
    un = base * (1 + 0.1 * np.random.default_rng(1234).uniform(-1, 1, N))
    unm1 = base * (1 + 0.1 * np.random.default_rng(1235).uniform(-1, 1, N))
    return un, unm1


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons}


def recorded_traffic(level, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    (profiles/r01_traffic.json); None when no capture exists for this workload / kernel."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        return t[str(level)][kernel]
    except Exception:
        return None


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's restatement of one time step on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------
# cylinder mesh size of the CPU sample mesh (0.025: 52 704 tets, 236 789 DoFs); the override exists for the tests
SAMPLE_LC = float(os.environ.get("NSB_BENCH_SAMPLE_LC", "0.025"))


def cpu_step_sample(repeats=1):
    """One pass of the hot path on the host cores with the oracle's C/OpenMP port (oracle/c/ns_oracle_c.c):
    the reference's assembly loops, then its solve_linear_system -- ILU(1) on F / ILU(0) on M_p over one row
    block per thread (Ifpack's per-rank ILU), CG-ILU(0) for K_p (in place of Trilinos ML), all re-factorized
    every solve like the reference, GMRES(150) to 1e-2*||b||, <= 200 iterations -- on a coarser mesh of the
    same geometry with the same synthetic state.  (On the mesh-3D-5-equivalent and finer the reference's
    ILU(1) of the grad-div-dominated F is unstable and GMRES stalls, so the sample is the finest mesh
    it converges on and is scaled linearly by the cell count, which favours the CPU.)
    Returns (cells, seconds per step, GMRES iterations, converged, threads)."""
    from oracle import assemble as asm, dofs as odofs, postprocess as pp, c_port
    from tools import meshgen
    mesh = meshgen.mesh_3d(lc_cyl=SAMPLE_LC, lc_global=0.15)
    dm = odofs.enumerate_dofs(mesh)
    pat = odofs.make_sparsity_fast(dm)
    tc = pp.TEST_CASES[CASE]
    con = odofs.build_constraints(mesh, dm, pp.inlet_profile(3, tc["U_m"], False, 4.0, 1.0), pp.boundary_ids(3))
    un, unm1 = synthetic_state(dm.support_points, dm.component, dm.n_u)
    p = asm.Params(dt=0.01, theta=0.5, nu=1e-3, use_supg=True)
    secs = []
    for _ in range(repeats):
        t0 = time.time()
        A, b, Mp, Kp = c_port.assemble_linearized(mesh, dm, pat, p, con, un, unm1, with_pressure_matrices=True)
        t1 = time.time()
        _, its, _, ok = c_port.solve(pat, dm.n_dofs, dm.n_u, A, Mp, Kp, b, p, max_it=200, tol_rel=1e-2, n_tmp_vectors=150)
        secs.append(time.time() - t0)
        log("[bench] cpu sample: assembly %.2f s, solve %.2f s, %d its, converged %s" % (t1 - t0, secs[-1] - (t1 - t0), its, ok))
    return mesh.n_cells, secs, its, ok, c_port.num_threads()


def cpu_baseline_entry(cells, sec, its, ok, threads, full_cells):
    v = (1.0 / sec) * (cells / full_cells)
    return {"value": v, "unit": "steps/s", "cores": threads, "kind": "port",
            "sample": "oracle C/OpenMP port (reference assembly loops; ILU(1)/ILU(0) per thread block + GMRES(150), re-factorized "
                      "every solve) on a %d-cell mesh of the same geometry and state: %.1f s per step, %d GMRES its, converged=%s; "
                      "scaled linearly by cells to the %d-cell workload" % (cells, sec, its, ok, full_cells)}


def reference_arm(args, full_cells):
    """`--impl reference`: the reference's own CPU path cannot be built here (deal.II + Trilinos + MPI
    are absent), so this times the oracle's C/OpenMP port on all host cores, as the task's tier rules prescribe.
    Each step is the bounded sample of cpu_step_sample()."""
    n = max(1, min(args.steps, 3))
    w = 1 if args.warmup else 0
    cells, secs, its, ok, threads = cpu_step_sample(repeats=n + w)
    sec = float(np.mean(secs[w:]))
    entry = cpu_baseline_entry(cells, sec, its, ok, threads, full_cells)
    value = entry["value"]
    line = {
        "impl": "reference", "metric": "time-steps/s", "value": value, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "mesh-3D-%d-equivalent, 3D-2Z (CN, linearised, SUPG+grad-div), GMRES(150) tol 1e-2" % args.level,
                   "timed_steps": n},
        "cpu_baseline": entry,
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--level", type=int, default=20, help="mesh-3D-<level>-equivalent (5, 10, 20, 40)")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--operator", type=int, default=0,
                    help="operator inside the velocity polynomial: 1 assembled fp32 copy, 2 element-wise, 0 library default")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    full_cells = {5: 92136, 10: 572922, 20: 3851622}.get(args.level, 3851622)

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, full_cells)
        return

    # Everything but the final JSON line goes to stderr -- including what C libraries (NCCL's version
    # banner) write to file descriptor 1.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    from tests.conftest import load_nsb
    nsb = load_nsb()
    dist = None
    uid = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        torch.cuda.set_device(local_rank)
        holder = [nsb.Device.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(holder, src=0)
        uid = holder[0]
    torch.cuda.set_device(local_rank)

    # ---- setup (not timed): mesh, DoFs, constraints on the host; structure upload
    if rank == 0:
        mesh_file = get_mesh_file(args.level)
    if world > 1:
        dist.barrier()
    mesh_file = get_mesh_file(args.level)
    t0 = time.time()
    hs = nsb.HostSetup(mesh_file, 3)
    pts, cells = hs.mesh()
    cell_dofs = hs.cell_dofs()
    sp_pts, comp = hs.support_points()
    cdofs, cvals = hs.constraints(CASE, 1.0)
    n_u, n_p, N = hs.n_u, hs.n_p, hs.n_dofs
    log("[bench] rank %d host setup %.1f s: %d cells, %d + %d DoFs" % (rank, time.time() - t0, hs.n_cells, n_u, n_p))
    dev = nsb.Device(3, local_rank)
    if args.operator:
        dev.set_solver_opts(precond_operator=args.operator)
    part = None
    if world > 1:
        dev.comm_init(rank, world, uid)
        part = (np.arange(hs.n_cells, dtype=np.int64) * world // hs.n_cells).astype(np.int32)
    t0 = time.time()
    dev.upload_mesh(pts, cells, cell_dofs, n_u, n_p, part)
    nrows, nnz, nc = dev.sizes()
    log("[bench] rank %d structure + upload %.1f s: %d rows, %d nnz" % (rank, time.time() - t0, nrows, nnz))
    un, unm1 = synthetic_state(sp_pts, comp, n_u)
    dev.set_constraints(cdofs, cvals)
    dev.set_params(0.01, 0.5, 1e-3, 1.0, 0.1, True, False)
    dev.set_vector(nsb.NSB_SOLUTION_OLD, un)
    dev.set_vector(nsb.NSB_SOLUTION_OLD_OLD, unm1)
    dev.assemble_linearized()
    t0 = time.time()
    dev.assemble_pressure_matrices()
    log("[bench] pressure matrices + multigrid setup %.1f s" % (time.time() - t0))

    def sync_all():
        dev.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident():
        dev.assemble_linearized()
        return dev.solve(200, 1e-2, 150)

    def step_e2e():
        dev.set_vector(nsb.NSB_SOLUTION_OLD, un)
        dev.set_vector(nsb.NSB_SOLUTION_OLD_OLD, unm1)
        dev.set_constraints(cdofs, cvals)
        dev.assemble_linearized()
        r = dev.solve(200, 1e-2, 150)
        dev.get_vector(nsb.NSB_SOLUTION, sol_host)
        return r

    sol_host = np.zeros(N)
    for _ in range(args.warmup):
        ok, its, res = step_resident()
    log("[bench] warm-up done: converged %s, %d GMRES iterations, %s" % (ok, its, dev.solver_info()))

    # ---- timed region 1: inputs resident in HBM, per-kernel CUDA-event profile on the library's stream
    sampler = ClockSampler(local_rank)
    sampler.start()
    dev.profile_enable(True)
    dev.profile_reset()
    launches0 = dev.launch_count()
    sync_all()
    dev.timer_start()
    iters = []
    for _ in range(args.steps):
        ok, its, res = step_resident()
        iters.append(its)
    ms = dev.timer_stop()
    sync_all()
    launches = dev.launch_count() - launches0
    log("[bench] rank %d timed region done: %.1f ms for %d steps" % (rank, ms, args.steps))
    prof = dev.profile()
    dev.profile_enable(False)
    t_res = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
    ms = float(t_res.item())
    # ---- timed region 2: end to end through the C ABI with host buffers
    step_e2e()
    sync_all()
    dev.timer_start()
    for _ in range(args.steps):
        step_e2e()
    ms_e2e = dev.timer_stop()
    sync_all()
    log("[bench] rank %d e2e region done: %.1f ms" % (rank, ms_e2e))
    t_e = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    ms_e2e = float(t_e.item())
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant kernel: the velocity-block SpMV (polynomial preconditioner).  Algorithmic bytes per
        # 12 nnz + 16 n + 4 (n+1), on the block it multiplies (F = A(0,0)).
        tot = {k: v[0] for k, v in prof.items()}
        dom = max(("spmv_vel", "spmv", "asm_rows", "orth"), key=lambda k: tot.get(k, 0.0))
        blocks = dev.block_nnz()
        opts_eff = dev.get_solver_opts()
        bytes_alg = {
            "spmv": 12 * nnz + 16 * nrows + 4 * (nrows + 1),
            "spmv_vel": None,
            "asm_rows": 8 * nnz + 8 * nrows + 8 * 3 * 4 * nc + 20 * 34 * nc,
        }
        kernels = {}
        for k, (t_ms, n) in prof.items():
            if n:
                kernels[k] = {"ms_total": round(t_ms, 3), "launches": n, "ms_avg": round(t_ms / n, 4)}
        line = {
            "metric": "time-steps/s", "value": args.steps / (ms * 1e-3), "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "mesh-3D-%d-equivalent (%d tets, %d DoFs, %d nnz), 3D-2Z: CN theta=0.5, linearised, dt=0.01, "
                                   "SUPG+grad-div, GMRES(150) tol 1e-2*||b||, max 200 its" % (args.level, hs.n_cells, N, nnz),
                       "parallelism": "1 process per GPU, contiguous cell chunks" if world > 1 else "single GPU",
                       "l2": "inputs (%.1f GB of matrix values) exceed L2; no flush needed" % (8e-9 * nnz),
                       "gmres_iterations_per_step": iters,
                       "solver": dict(dev.solver_info(), velocity_operator={1: "assembled fp%d copy" % opts_eff["precond_precision"],
                                                                            2: "element-wise (S rows fp32 + cell geometry)"}[opts_eff["precond_operator"]])},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": int(2 * 8 * N + 12 * cdofs.size),
                    "d2h_bytes_per_step": int(8 * N)},
            "kernels": kernels,
        }
        # roofline of the dominant kernel
        bytes_alg["spmv_vel"] = 12 * blocks["uu"] + 16 * n_u + 4 * (n_u + 1)
        if dom in kernels and bytes_alg.get(dom):
            ach = bytes_alg[dom] / (kernels[dom]["ms_avg"] * 1e-3) / 1e9
            line["roofline"] = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                "traffic": recorded_traffic(args.level, dom + ("_ebe" if dom == "spmv_vel" and opts_eff["precond_operator"] == 2 else "")) if world == 1 else None, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(bytes_alg[dom])}
            tr = line["roofline"]["traffic"]
            if tr:
                # what the kernel really moves (compressed indices, fp32 operator copy) against the same peak
                line["roofline"]["dram_GBps"] = tr / (kernels[dom]["ms_avg"] * 1e-3) / 1e9
                line["roofline"]["dram_frac"] = line["roofline"]["dram_GBps"] / peak
        for k in ("spmv", "asm_rows"):
            if k in kernels:
                kernels[k]["algorithmic_GBps"] = round(bytes_alg[k] / (kernels[k]["ms_avg"] * 1e-3) / 1e9, 1)
        if not args.no_cpu_baseline and world == 1:
            cells_s, secs, its_s, ok_s, thr = cpu_step_sample(repeats=3)      # first pass warms the caches / thread pool
            line["cpu_baseline"] = cpu_baseline_entry(cells_s, float(np.mean(secs[1:])), its_s, ok_s, thr, hs.n_cells)
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    dev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
