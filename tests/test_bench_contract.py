"""bench.py's reference arm (CPU only): one JSON line on stdout with the keys the driver parses, measured with the
oracle's C/OpenMP port on the SAME workload the GPU arm names (here the mesh-3D-5-equivalent, GMRES cut short by the arm's
time budget to keep the test fast); all host cores even under torchrun's OMP_NUM_THREADS=1; non-zero ranks print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    env = dict(os.environ, NSB_BENCH_REF_BUDGET_S="3", OMP_NUM_THREADS="1", **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--level", "5", "--steps", "20", "--warmup", "5"],
                          capture_output=True, text=True, env=env, timeout=900)


def test_reference_arm_prints_one_contract_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "time-steps/s" and d["unit"] == "steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    # what was really timed: one step, no warm-up -- ms_per_step * steps must fit inside the run
    assert d["steps"] == 1 and d["warmup"] == 0 and d["config"]["requested_steps"] == 20
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["same_config"] is True
    assert cb["cores"] == len(os.sched_getaffinity(0))                 # not torchrun's OMP_NUM_THREADS=1
    if len(os.sched_getaffinity(0)) > 1:
        assert cb["cores"] > 1
    assert "scaled" not in cb["sample"] and "mesh-3D-5-equivalent" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the workload string is the GPU arm's (bench.workload_name), sizes included
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"].startswith("mesh-3D-5-equivalent (")
    assert d["config"]["workload"].endswith(bench.workload_name(5).split("equivalent", 1)[1])
    bd = d["config"]["breakdown_s"]
    assert bd["assembly_s"] > 0 and bd["precond_setup_s"] > 0 and bd["gmres_iterations"] >= 1


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
