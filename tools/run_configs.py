"""Runs BASELINE.json's smaller configurations through the C++ driver (the drop-in executable that plays the
role of the reference's src/main.cpp) and keeps forces.txt + the log:  python -m tools.run_configs OUTDIR"""
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import meshgen, msh  # noqa: E402

EXE = os.path.join(ROOT, "navier-stokes_equations_b200", "navier_stokes")


def run(case, mesh_path, steps, out_dir, tag, extra=()):
    work = os.path.join("/tmp", "nsb_run_" + tag)
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(work)
    t0 = time.time()
    r = subprocess.run([EXE, case, mesh_path, "--steps", str(steps)] + list(extra), cwd=work, capture_output=True, text=True)
    dt = time.time() - t0
    open(os.path.join(out_dir, tag + ".log"), "w").write(r.stdout[-20000:] + "\n--- stderr ---\n" + r.stderr[-4000:])
    if os.path.exists(os.path.join(work, "forces.txt")):
        shutil.copy(os.path.join(work, "forces.txt"), os.path.join(out_dir, tag + "_forces.txt"))
    vt = [f for f in os.listdir(work) if f.endswith(".vtu") or f.endswith(".pvtu")]
    print(f"{tag}: rc={r.returncode} {dt:.1f}s for {steps} steps, vtu files {len(vt)}", flush=True)
    its = [ln for ln in r.stdout.splitlines() if "GMRES" in ln]
    print("   ", its[:3], "...", its[-2:], flush=True)
    print("   ", [ln for ln in r.stdout.splitlines() if "Cd=" in ln][-1:], flush=True)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
    os.makedirs(out, exist_ok=True)
    # config 1: mesh-2D (shipped), 2D-1 Re 20 BE + Newton
    m = msh.load_npz(os.path.join(ROOT, "tests/golden/mesh-2D.npz"))
    msh.write_msh("/tmp/mesh-2D.msh", m)
    run("2D-1", "/tmp/mesh-2D.msh", 5, out, "config1_2D-1_mesh-2D", ["--no-vtu"])
    # config 2: mesh-2D-200-equivalent (red refinement of the shipped mesh-2D-100), 2D-2 Re 100 CN + linearised
    m200 = meshgen.refine_2d(msh.load_npz(os.path.join(ROOT, "tests/golden/mesh-2D-100.npz")))
    msh.write_msh("/tmp/mesh-2D-200.msh", m200)
    print("mesh-2D-200-equivalent:", m200.n_cells, "cells", flush=True)
    run("2D-2", "/tmp/mesh-2D-200.msh", 40, out, "config2_2D-2_mesh-2D-200eq", ["--no-vtu"])
    # config 3: mesh-3D-5-equivalent, 3D-2Z; the first two steps also write VTU
    m5 = meshgen.mesh_3d(5)
    msh.write_bin("/tmp/mesh-3D-5.bin", m5)
    run("3D-2Z", "/tmp/mesh-3D-5.bin", 20, out, "config3_3D-2Z_mesh-3D-5eq", ["--no-vtu"])
    run("3D-2Z", "/tmp/mesh-3D-5.bin", 1, out, "config3_vtu_check", [])


if __name__ == "__main__":
    main()
