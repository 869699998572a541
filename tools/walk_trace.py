"""Where a launch of the landmark half spends its time (tuning build: python -c "from povar_b200 import build;
build.build(defines=['POVAR_WALK_TRACE'], suffix='_trace')", POVAR_LIB=povar_b200/lib/libpovar_b200_trace.so):
per block the globaltimer stamps [entry, window staged, slices walked, done] of the last launch.

    POVAR_LIB=povar_b200/lib/libpovar_b200_trace.so python tools/walk_trace.py [workload] [shard_of] [pose|joint]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from povar_b200 import capi, synthetic  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "venice1778"
shard_of = int(sys.argv[2]) if len(sys.argv) > 2 else 1
step = sys.argv[3] if len(sys.argv) > 3 else "pose"
sp = synthetic.generate_named(workload)
hp = capi.HostProblem.from_unordered(sp.num_cams, sp.num_lms, sp.obs_cam, sp.obs_lm, sp.obs_xy, sp.cam_params)
if shard_of > 1:
    hp = hp.shard(0, shard_of)
opt = capi.default_options(alpha=0.1, power_sc_iterations=20, verbosity_level=0, robust_norm=capi.NORM_CAUCHY)
s = capi.Solver(hp, opt)
s.initialize_varproj_lm_pOSE(0.1)
s.linearize_pOSE(0.1)
s.solve(1e-4)
which = capi.STATE_POSE
if step == "joint":
    s.backup(capi.STATE_POSE)
    s.apply(0.1)
    s.to_homogeneous()
    s.linearize_projective_space_homogeneous()
    s.solve_joint(1e-4)
    which = capi.STATE_JOINT
lib = capi.load()
buf = (C.c_uint64 * 4096)()
lib.povar_debug_walk_trace(buf, 4096)          # reset
k = s.bench_power_kernels(which, 1)              # one launch of each kernel of a term; the landmark half comes first
rc = lib.povar_debug_walk_trace(buf, 4096)
assert rc == 0, rc
t = np.array(buf, dtype=np.uint64).reshape(-1, 4).astype(np.float64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
t = (t - t0) / 1e3
print(f"{workload} 1/{shard_of} {step}: {len(t)} blocks, kernel {1e6 * float(k[0]):.1f} us (events)")
for i, name in enumerate(("entry", "window staged", "slices walked", "done")):
    print(f"  {name:14s} min {t[:, i].min():7.2f}  median {np.median(t[:, i]):7.2f}  max {t[:, i].max():7.2f} us")
s.close()
