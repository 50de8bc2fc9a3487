"""Failure paths of run() (reference src/classes/NavierStokes.cpp:1174-1198 Newton `(linfail)` + backtracking,
:1223-1286 linearised dt-halving retry with checkpoint restore, Backward-Euler fallback, forced BE) and the
output files (:1013-1042, :1064-1068, :1315-1319), driven through the C++ host class on the GPU and compared
with the oracle's restatement of the same control flow under the same forced failures.

The failures are forced with the `test_fail_solves` hook: the next k linear solves are solved normally (tight
tolerance, so both sides hold the exact solution of each system) but REPORTED as not converged -- which is what
selects the branch; the iterate of a "failed" solve is kept, as in the reference."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from oracle import solve as osolve
from tests.conftest import GOLDEN  # noqa: F401

pytestmark = pytest.mark.gpu

TOL_FORCE = 1e-6
TOL_FIELD = 1e-8


def _close(info, ref, k):
    for key in ("cd", "cl", "dp"):
        assert abs(info[key] - ref[key]) <= TOL_FORCE * abs(ref[key]) + 1e-12, (k, key, info[key], ref[key])


# k forced failures -> (linear solves executed in that step, step accepted?)   cpp:1223-1286
#   1: attempt 0 fails, BE + first-order fallback accepted
#   2: both fail, retry with dt/2 accepted
#   4: ... retry with dt/8 accepted
#   6: attempt 0, fallback and the four halvings fail -> checkpoint restored, forced BE at dt/16 kept
@pytest.mark.parametrize("k,solves,accepted", [(1, 2, True), (2, 3, True), (4, 5, True), (6, 7, False)])
def test_linearized_retry_sequence_matches_oracle(nsb, msh_file, golden_mesh, capfd, k, solves, accepted):
    s = nsb.HostSolver("2D-2", msh_file("mesh-2D"), gmres_tolerance=1e-12, verbose=True)
    s.initialize()
    o = osolve.Oracle(golden_mesh("mesh-2D"), "2D-2", solver="direct")
    for _ in range(2):                       # BE first step, second step: no failures
        info, ref = s.step(), o.step()
        _close(info, ref, "pre")
    capfd.readouterr()
    s.set_test_fail_solves(k)
    o.fail_solves = k
    t_before = info["time"]
    info, ref = s.step(), o.step()
    out = capfd.readouterr().out
    assert info["solves"] == solves == len(ref["gmres"])
    assert bool(info["converged"]) == accepted == ref["ok"]
    assert abs(info["time"] - (t_before + o.deltat)) < 1e-14          # `time` advances by the FULL dt (cpp:1074)
    _close(info, ref, k)
    x = s.solution()
    assert np.linalg.norm(x - o.current_solution) / np.linalg.norm(o.current_solution) < TOL_FIELD
    # the reference's messages, in its order
    assert "Fallback to BE + 1st-order..." in out
    n_retry = out.count("Retrying with dt=")
    assert n_retry == max(0, min(k - 1, 4))
    if k >= 2 and accepted:
        assert "Step accepted with reduced dt=%g" % (o.deltat / 2 ** (k - 1)) in out
    if not accepted:
        assert "CRITICAL: all attempts failed. Restoring checkpoint and forcing BE dt=%g" % (o.deltat / 16) in out
        assert out.index("Fallback to BE") < out.index("Retrying with dt=") < out.index("CRITICAL")
    # the state the failure leaves behind is the reference's: one more (clean) step still agrees
    info, ref = s.step(), o.step()
    assert info["solves"] == 1 and info["converged"] == 1
    _close(info, ref, "post")
    s.close()


def test_newton_linfail_damping_matches_oracle(nsb, msh_file, golden_mesh, capfd):
    """First Newton iteration of 2D-1 reported as failed: `(linfail)`, damping x0.25, re-assembly and the
    backtracking test (cpp:1174-1198); the iteration then recovers.  Same iteration count and forces as the
    oracle under the same failure."""
    s = nsb.HostSolver("2D-1", msh_file("mesh-2D"), gmres_tolerance=1e-12, verbose=True)
    s.initialize()
    o = osolve.Oracle(golden_mesh("mesh-2D"), "2D-1", solver="direct")
    s.set_test_fail_solves(1)
    o.fail_solves = 1
    capfd.readouterr()
    info, ref = s.step(), o.step()
    out = capfd.readouterr().out
    assert "(linfail)" in out and " a=0.25" in out
    assert info["newton_iterations"] == ref["newton_iters"] > 2
    _close(info, ref, "newton")
    x = s.solution()
    assert np.linalg.norm(x - o.current_solution) / np.linalg.norm(o.current_solution) < TOL_FIELD
    s.close()


def _parse_piece(path):
    root = ET.parse(path).getroot()
    piece = root.find("UnstructuredGrid/Piece")
    arrays = {}
    for da in piece.iter("DataArray"):
        arrays[da.get("Name") or "points"] = np.array(da.text.split(), dtype=float)
    return int(piece.get("NumberOfPoints")), int(piece.get("NumberOfCells")), arrays


def test_output_files(nsb, msh_file, golden_mesh, tmp_path):
    """forces.txt (header + one 6-significant-digit row per step, cpp:1064-1068, 1315-1319) and
    solution_NNNN.pvtu / .vtu with velocity, pressure, subdomain (cpp:1013-1042)."""
    out = str(tmp_path) + "/"
    m = golden_mesh("mesh-2D")
    s = nsb.HostSolver("2D-2", msh_file("mesh-2D"), write_vtu=True, output_dir=out)
    s.initialize()
    infos = [s.step() for _ in range(2)]
    x = s.solution()
    s.close()
    rows = open(out + "forces.txt").read().strip().split("\n")
    assert rows[0] == "Time\tCd\tCl\tDeltaP" and len(rows) == 3
    for r, info in zip(rows[1:], infos):
        got = [float(v) for v in r.split("\t")]
        ref = [info["time"], info["cd"], info["cl"], info["dp"]]
        assert got == [float("%.6g" % v) for v in ref]            # default ostream precision: 6 significant digits
    for step in range(3):
        p = ET.parse(out + "solution_%04d.pvtu" % step).getroot()
        pieces = [e.get("Source") for e in p.iter("Piece")]
        assert pieces == ["solution_%04d.0.vtu" % step]
        names = {e.get("Name") for e in p.iter("PDataArray")}
        assert {"velocity", "pressure", "subdomain"} <= names
    npts, ncells, arr = _parse_piece(out + "solution_0002.0.vtu")
    assert npts == m.n_vertices and ncells == m.n_cells
    assert np.all(arr["types"] == 5) and np.all(arr["subdomain"] == 0)
    conn = arr["connectivity"].astype(int).reshape(-1, 3)
    pts = arr["points"].reshape(-1, 3)
    # same triangles (points are renumbered in order of first use)
    assert np.allclose(pts[conn][:, :, :2], m.points[m.cells], atol=1e-8)
    # vertex values of the written fields = the solution vector at the vertex DoFs
    from oracle import dofs as odofs
    dm = odofs.enumerate_dofs(m)
    vel = arr["velocity"].reshape(-1, 3)
    first_use = np.full(m.n_vertices, -1)
    for vtx, loc in zip(m.cells.ravel(), conn.ravel()):
        first_use[vtx] = loc
    vdofs = dm.cell_dofs[:, [0, 3, 6]]          # u_0 of the three vertices of each cell
    pdofs = dm.cell_dofs[:, [2, 5, 8]]
    assert np.allclose(vel[conn][:, :, 0], x[vdofs], rtol=1e-7, atol=1e-12)
    assert np.allclose(vel[conn][:, :, 1], x[vdofs + 1], rtol=1e-7, atol=1e-12)
    assert np.allclose(arr["pressure"][conn], x[pdofs], rtol=1e-7, atol=1e-12)
