"""Run the large golden configurations on the GPU and dump the traces (cost, accept, linear iterations)
to gpurun_out/large_traces.json for offline comparison with tests/golden/traces_large.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import povar_testlib as common  # noqa: E402
from povar_b200 import capi  # noqa: E402

out = {}
names = sys.argv[1:] or list(common.traces_large()["traces"])
for name in names:
    meta = common.traces_large()["traces"][name]
    kw = common.flags_to_options(meta["flags"])
    hp = capi.HostProblem.read(common.golden_file(meta["shape"]))
    s = capi.Solver(hp, capi.default_options(verbosity_level=0, **kw))
    t = time.time()
    its, summary = s.bundle_adjust()
    dt = time.time() - t
    s.close()
    out[name] = {"cost": [e.cost for e in its], "succ": [int(e.step_is_successful) for e in its],
                 "lin": [e.linear_solver_iterations for e in its], "iteration": [e.iteration for e in its],
                 "tr": [e.trust_region_radius for e in its], "wall_s": dt, "solve_s": summary.total_time}
    print(name, len(its), "trials", f"{dt:.2f}s", "final", its[-1].cost, "ref", meta["threads1"]["cost"][-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "large_traces.json"), "w") as f:
    json.dump(out, f)
