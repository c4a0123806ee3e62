"""bench.py contract (CPU part): the reference arm prints ONE JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=600, env=e)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_reference_arm_json_line():
    so = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0", "--workload", "24x20"])
    lines = [ln for ln in so.splitlines() if ln.strip()]
    assert len(lines) == 1, so                      # exactly one line on stdout (libraries print to stderr)
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "IMPES cell-steps/s" and d["unit"] == "cell-steps/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["higher_is_better"] is True
    assert d["value"] > 0 and abs(d["value"] - 24 * 20 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["config"]["workload"] == "24x20" and len(d["iters_per_step"]) == 1
    # the config object is a pure function of the command line (identical in both arms): no measured values in it
    assert set(d["config"]) == {"workload", "mesh", "spaces", "two_phase", "K", "sat_source", "dt_global", "dt_local",
                                "substeps_per_step", "schur_prec", "minres_rtol", "minres_maxiter", "partition",
                                "l2_policy"}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    """Under torchrun (N > 1) rank 0 alone runs and prints the reference arm; the other ranks exit 0 without work."""
    so = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "12x10"],
              env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert so.strip() == ""


def test_reference_arm_uses_all_cores_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must still use every host core."""
    so = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "12x10"],
              env={"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0", "OMP_NUM_THREADS": "1"})
    d = json.loads(so.strip())
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_default_workload_is_c5_at_every_gpu_count():
    sys.path.insert(0, ROOT)
    import importlib
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'name = "C5"' in src and "gpus == 1" not in src
    from __graft_entry__ import load_package
    load_package()
    from incompressible_flow_b200 import workloads
    c, dtg, dtl = workloads.bench_setup("C5")
    assert (c["nx"], c["ny"], c["nz"], c["order"]) == (256, 256, 256, 1) and c.get("order_s", 1) == 1
    assert dtg / dtl == 100 and c["saturation_source"] == 0.5
