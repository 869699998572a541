"""Set up the venice-1778-shaped problem and run a few power-series terms: the command ncu wraps.

    ncu --set full --clock-control none --import-source on -k regex:k_e0_landmark -c 2 -o gpurun_out/prof \
        python tools/profile_term.py [workload] [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from povar_b200 import capi, synthetic  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "venice1778"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
step = sys.argv[3] if len(sys.argv) > 3 else "pose"
shard_of = int(sys.argv[4]) if len(sys.argv) > 4 else 1     # > 1: the first of that many landmark shards, alone
sp = synthetic.generate_named(workload)
hp = capi.HostProblem.from_unordered(sp.num_cams, sp.num_lms, sp.obs_cam, sp.obs_lm, sp.obs_xy, sp.cam_params)
if shard_of > 1:
    hp = hp.shard(0, shard_of)
opt = capi.default_options(alpha=0.1, power_sc_iterations=20, verbosity_level=0, robust_norm=capi.NORM_CAUCHY,
                           max_num_iterations_step_1=3, max_num_iterations_step_2=2)
s = capi.Solver(hp, opt)
if step == "solve":
    its, summ = s.bundle_adjust()
    print("solve", summ.total_time, len(its))
else:
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
    k = s.bench_power_kernels(which, reps)
    print("kernel us:", [round(1e6 * float(v), 1) for v in k], "| term in sequence us:",
          round(1e6 * s.bench_power_terms(which, 2 * reps), 1))
s.close()
