"""Per-function parity checks: CUDA path (through the C ABI) vs the numpy oracle.

Used by tests/test_gpu_parity.py; `python tests/povar_gpu_checks.py [shape ...]` prints the same
numbers as a report (needs a GPU; run under gpurun).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import povar_oracle as O  # noqa: E402
from povar_b200 import capi, synthetic  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / den


def sym6_to_33(h):
    out = np.empty((h.shape[0], 3, 3))
    idx = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    for k, (i, j) in enumerate(idx):
        out[:, i, j] = h[:, k]
        out[:, j, i] = h[:, k]
    return out


def make(shape, seed=None, **kw):
    sp = synthetic.generate_named(shape, seed, **kw)
    hp = capi.HostProblem.from_unordered(sp.num_cams, sp.num_lms, sp.obs_cam, sp.obs_lm, sp.obs_xy, sp.cam_params)
    op = O.build_problem(sp.num_cams, sp.num_lms, sp.obs_cam.astype(np.int64), sp.obs_lm.astype(np.int64),
                         sp.obs_xy, sp.cam_params)
    return sp, hp, op


def check_shape(shape, step1=capi.POWER_VARPROJ, robust=capi.NORM_NONE, alpha=0.1, m=20, lam1=1e-4,
                lam2=1e-4, huber=1.0, report=print):
    sp, hp, op = make(shape)
    C, L = hp.num_cams, hp.num_lms
    report(f"== {shape}: {C} cams, {L} lms, {hp.num_obs} obs; step1 solver {step1}, robust {robust}")
    assert np.array_equal(hp.lm_ptr, op.lm_ptr) and np.array_equal(hp.obs_cam, op.obs_cam)
    assert np.array_equal(hp.obs_uv, op.uv)
    opt = capi.default_options(alpha=alpha, power_sc_iterations=m, solver_type_step_1=step1, robust_norm=robust,
                               huber_parameter=huber, verbosity_level=0)
    oopt = O.Options(alpha=alpha, power_sc_iterations=m, solver_type_step_1=step1, robust_norm=robust,
                     huber_parameter=huber)
    s = capi.Solver(hp, opt)
    worst = {}

    def rec(name, val):
        worst[name] = val
        report(f"   {name:28s} {val:.3e}")

    # ---- step 1
    s.initialize_varproj_lm_pOSE(alpha)
    O.init_varproj(op, alpha)
    P, X = s.get_state(capi.STATE_POSE)
    rec("init_varproj X", rel(X, op.X))
    ri = s.compute_error_pOSE(alpha)
    ori = O.cost_pose(op, oopt)
    rec("cost_pose", abs(ri.error_all - ori.err_all) / ori.err_all)
    # use the oracle's landmarks from here on so later stages are compared on identical inputs
    s.set_state(capi.STATE_POSE, None, op.X)

    lz = O.PoseLinearizor(op, oopt)
    lz.linearize()
    assert s.linearize_pOSE(alpha) == capi.OK
    rec("pose_scale", rel(s.debug_read("pose_scale").reshape(C, 12), lz.lin.pose_scale))
    rec("lm_scale", rel(s.debug_read("lm_scale").reshape(L, 4)[:, :3], lz.lin.jl_scale))
    lam = lam1
    oinc, oit = lz.solve(lam)
    inc, it, rc = s.solve(lam)
    rec("hll_inv", rel(sym6_to_33(s.debug_read("hll_inv").reshape(L, 6)), lz.dbg["Hll_inv"]))
    rec("b", rel(s.debug_read("b").reshape(C, 12), lz.dbg["b"]))
    rec("b_inv", rel(s.debug_read("b_inv").reshape(C, 12, 12), lz.dbg["B_inv"]))
    x = np.random.default_rng(5).normal(size=(C, 12))
    e0 = s.right_mul_e0(capi.STATE_POSE, x)
    oe0 = O.right_mul_e0(lz.lin.Jp, lz.lin.Jl, lz.dbg["Hll_inv"], op, x)
    rec("right_mul_e0", rel(e0, oe0))
    rec("inc (power series)", rel(inc, oinc))
    report(f"   linear_solver_iterations      {it} vs {oit}; status {rc}")
    worst["lin_it_equal"] = 0.0 if it == oit else 1.0
    worst["lin_it"] = it
    s.backup(capi.STATE_POSE)
    l_diff = s.apply(alpha)
    ol = lz.apply(oinc)
    rec("l_diff", abs(l_diff - ol) / abs(ol))
    P, X = s.get_state(capi.STATE_POSE)
    rec("apply P", rel(P, op.P))
    rec("apply X", rel(X, op.X))
    ri = s.compute_error_pOSE(alpha)
    ori = O.cost_pose(op, oopt)
    rec("cost_pose after step", abs(ri.error_all - ori.err_all) / ori.err_all)

    # ---- step 2 from the oracle's state
    s.set_state(capi.STATE_POSE, op.P, op.X)
    s.to_homogeneous()
    O.to_homogeneous(op)
    P, Xh = s.get_state(capi.STATE_JOINT)
    rec("to_homogeneous P", rel(P, op.P))
    rec("to_homogeneous X", rel(Xh, op.Xh))
    ri = s.compute_error_homogeneous()
    ori = O.cost_joint(op, oopt)
    rec("cost_joint", abs(ri.error_all - ori.err_all) / ori.err_all)
    worst["valid_equal"] = 0.0 if ri.num_obs_valid == ori.n_valid else 1.0
    jz = O.JointLinearizor(op, oopt)
    jz.linearize()
    assert s.linearize_projective_space_homogeneous() == capi.OK
    rec("joint pose_scale", rel(s.debug_read("pose_scale").reshape(C, 12), jz.lin.pose_scale))
    rec("joint lm_scale", rel(s.debug_read("lm_scale").reshape(L, 4), jz.lin.jl_scale))
    lam = lam2
    oinc, oit = jz.solve(lam)
    inc, it, rc = s.solve_joint(lam)
    rec("joint hll_inv", rel(sym6_to_33(s.debug_read("hll_inv").reshape(L, 6)), jz.dbg["Hll_inv"]))
    rec("joint b", rel(s.debug_read("b").reshape(C, 11), jz.dbg["b"]))
    rec("joint b_inv", rel(s.debug_read("b_inv").reshape(C, 144)[:, :121].reshape(C, 11, 11), jz.dbg["B_inv"]))
    x = np.random.default_rng(6).normal(size=(C, 11))
    e0 = s.right_mul_e0(capi.STATE_JOINT, x)
    oe0 = O.right_mul_e0(jz.lin.Jp_t, jz.lin.Jl_t, jz.dbg["Hll_inv"], op, x)
    rec("joint right_mul_e0", rel(e0, oe0))
    rec("joint inc", rel(inc, oinc))
    report(f"   linear_solver_iterations      {it} vs {oit}; status {rc}")
    s.backup(capi.STATE_JOINT)
    l_diff = s.apply_joint()
    ol = jz.apply(oinc)
    rec("joint l_diff", abs(l_diff - ol) / abs(ol))
    s.normalize_joint()
    O.normalize_joint(op)
    P, Xh = s.get_state(capi.STATE_JOINT)
    rec("joint apply P", rel(P, op.P))
    rec("joint apply X", rel(Xh, op.Xh))
    ri = s.compute_error_homogeneous()
    ori = O.cost_joint(op, oopt)
    rec("cost_joint after step", abs(ri.error_all - ori.err_all) / ori.err_all)
    s.close()
    return worst


def check_trace(shape, report=print, **optkw):
    sp, hp, op = make(shape)
    kw = dict(alpha=0.1, power_sc_iterations=20)
    kw.update(optkw)
    opt = capi.default_options(verbosity_level=0, **kw)
    oopt = O.Options(**kw)
    s = capi.Solver(hp, opt)
    t = time.time()
    its, summary = s.bundle_adjust()
    t_gpu = time.time() - t
    t = time.time()
    olog = O.bundle_adjust(op, oopt)
    t_cpu = time.time() - t
    report(f"== trace {shape} {optkw}: gpu {len(its)} its in {t_gpu:.2f}s (solve {summary.total_time:.3f}s), "
           f"oracle {len(olog)} its in {t_cpu:.1f}s")
    n = min(len(its), len(olog))
    worst = 0.0
    first_bad = None
    for i in range(n):
        a, b = its[i], olog[i]
        r = abs(a.cost - b.cost) / abs(b.cost)
        same = (bool(a.step_is_successful) == b.step_is_successful and
                a.linear_solver_iterations == b.linear_solver_iterations)
        if first_bad is None and (r > 1e-9 or not same):
            first_bad = i
        worst = max(worst, r)
        if i < 6 or i >= n - 2 or not same:
            report(f"   {i:3d} step{a.step} it{a.iteration:2d} cost {a.cost:.12e} vs {b.cost:.12e} rel {r:.1e} "
                   f"succ {a.step_is_successful}/{int(b.step_is_successful)} lin {a.linear_solver_iterations}/"
                   f"{b.linear_solver_iterations}")
    report(f"   worst rel {worst:.2e}; first deviation at {first_bad}; final {its[-1].cost:.9e} vs {olog[-1].cost:.9e}")
    s.close()
    return worst, first_bad, len(its), len(olog)


if __name__ == "__main__":
    shapes = sys.argv[1:] or ["tiny", "small", "ladybug49"]
    for sh in shapes:
        check_shape(sh)
    check_shape("small", step1=capi.POWER_SCHUR_COMPLEMENT)
    check_shape("small", robust=capi.NORM_HUBER)
    check_trace("small")
    check_trace("small", solver_type_step_1=capi.POWER_SCHUR_COMPLEMENT)
