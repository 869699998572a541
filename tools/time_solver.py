"""Wall time of a (possibly capped) two-step solve of a synthetic shape with a given step-1 solver:
python tools/time_solver.py SHAPE SOLVER [max_it_step1 max_it_step2]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from povar_b200 import capi, synthetic  # noqa: E402

shape, solver = sys.argv[1], sys.argv[2]
kw = {}
if len(sys.argv) > 4:
    kw = {"max_num_iterations_step_1": int(sys.argv[3]), "max_num_iterations_step_2": int(sys.argv[4])}
sp = synthetic.generate_named(shape)
hp = capi.HostProblem.from_unordered(sp.num_cams, sp.num_lms, sp.obs_cam, sp.obs_lm, sp.obs_xy, sp.cam_params)
opt = capi.default_options(alpha=0.1, power_sc_iterations=20, verbosity_level=0,
                           solver_type_step_1=getattr(capi, solver), **kw)
s = capi.Solver(hp, opt)
t = time.perf_counter()
its, summ = s.bundle_adjust()
dt = time.perf_counter() - t
print(f"{shape} {solver}: {len(its)} trials, {dt:.3f} s, {1e3 * dt / max(len(its), 1):.1f} ms per trial, "
      f"final cost {its[-1].cost:.6e}")
s.close()
