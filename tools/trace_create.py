"""Phase times of povar_create / solve / read-back / destroy from host buffers, a few times in one process (the
first handle pays for the memory pool): POVAR_TRACE_CREATE=1 python tools/trace_create.py [workload] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("POVAR_TRACE_CREATE", "1")
from povar_b200 import capi, synthetic  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "venice1778"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sp = synthetic.generate_named(workload)
hp = capi.HostProblem.from_unordered(sp.num_cams, sp.num_lms, sp.obs_cam, sp.obs_lm, sp.obs_xy, sp.cam_params)
opt = capi.default_options(alpha=0.1, power_sc_iterations=20, verbosity_level=0, robust_norm=capi.NORM_CAUCHY)
# page-locked inputs and outputs, as bench.py's end-to-end leg has them
import numpy as np  # noqa: E402
import torch  # noqa: E402
rt = torch.cuda.cudart()
out = (np.zeros((hp.num_cams, 3, 4)), np.zeros((hp.num_lms, 4)))
for arr in (hp.lm_ptr, hp.obs_cam, hp.obs_uv, hp.cam_P) + out:
    rt.cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)
for r in range(reps):
    t0 = time.perf_counter()
    s = capi.Solver(hp, opt)
    t1 = time.perf_counter()
    its, summ = s.bundle_adjust()
    t2 = time.perf_counter()
    P, X = s.get_state(capi.STATE_JOINT, out=out)
    t3 = time.perf_counter()
    s.close()
    t4 = time.perf_counter()
    print(f"rep {r}: create {1e3 * (t1 - t0):.1f} ms, solve {1e3 * (t2 - t1):.1f} ms, read-back {1e3 * (t3 - t2):.1f} ms, "
          f"destroy {1e3 * (t4 - t3):.1f} ms", file=sys.stderr)
