"""The reference's own driver on the GPU library, timed: oracle/_ref/bal_ref_b200 (the reference program with its
linearizors replaced by integration/linearizor_b200.hpp, `make -C oracle plugin`) on the data_custom file of the
benchmark shape, the benchmark's flags, the full 50 + 50 iteration budget.  Prints one JSON line (profiles/
r2_plugin_bench.json); the times are the reference driver's own `ba_log.json` columns.

    python tools/plugin_bench.py [workload] [reps]
"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from povar_b200 import synthetic  # noqa: E402

PLUGIN = os.path.join(ROOT, "oracle", "_ref", "bal_ref_b200")
workload = sys.argv[1] if len(sys.argv) > 1 else "venice1778"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if not os.path.exists(PLUGIN):
    sys.exit("oracle/_ref/bal_ref_b200 not built (make -C oracle plugin)")
sp = synthetic.generate_named(workload)
runs = []
with tempfile.TemporaryDirectory() as work:
    path = os.path.join(work, "data_custom", f"{workload}.txt")
    os.makedirs(os.path.dirname(path))
    t = time.time()
    synthetic.write_bal(sp, path)
    t_write = time.time() - t
    for r in range(reps):
        log = os.path.join(work, f"ba_log_{r}.json")
        cmd = [PLUGIN, "--input", path, "--num-threads", str(os.cpu_count() or 1), "--alpha", "0.1",
               "--power-sc-iterations", "20", "--solver-type-step-1", "POWER_VARPROJ", "--solver-type-step-2", "RIPOBA",
               "--residual-robust-norm", "CAUCHY", "--log-log-path", log]
        t = time.time()
        res = subprocess.run(cmd, capture_output=True, text=True, cwd=work)
        wall = time.time() - t
        if res.returncode != 0:
            sys.exit("bal_ref_b200 failed: " + (res.stderr or res.stdout)[-1500:])
        with open(log) as f:
            d = json.load(f)
        timing = d["_static"]["timing"]
        runs.append({
            "wall_s": wall, "load_s": timing.get("load"), "optimize_s": timing.get("optimize"),
            "lm_trials": len(d["iteration"]), "final_cost": d["cost"][-1], "power_terms": sum(d["linear_solver_iterations"]),
            "phase_s": {k: sum(d[k]) for k in ("jacobian_evaluation_time", "prepare_time", "solve_reduced_system_time",
                                                "back_substitution_time", "residual_evaluation_time") if k in d},
        })
best = min(runs, key=lambda r: r["optimize_s"])
print(json.dumps({
    "what": "reference driver (bal_ref_b200) on libpovar_b200.so through LinearizorB200, one B200",
    "workload": workload, "cameras": sp.num_cams, "landmarks": sp.num_lms, "observations": sp.num_obs,
    "flags": "--alpha 0.1 --power-sc-iterations 20 POWER_VARPROJ + RIPOBA, CAUCHY, 50 + 50 iterations",
    "lm_iterations_per_s": best["lm_trials"] / best["optimize_s"], "optimize_s": best["optimize_s"],
    "runs": runs, "write_file_s": t_write,
}))
