"""Parity with the reference at BASELINE.json's sizes: full two-step traces of the CUDA path against
`bal_ref --num-threads 1` on the same (regenerated, hash-checked) data_custom files.

  c2  trafalgar-257 shape, all four step-1 solvers    c3  venice-89 shape, POWER_SCHUR_COMPLEMENT
  c4  venice-1778 shape, CAUCHY (the benchmark configuration)

What can be asked at these sizes is set by the reference itself: its own 8-thread run (stored next to the
1-thread run; only the order of its scatter-adds differs) agrees with the 1-thread run to 1e-11 through step 1
(1e-8 with PCG / CHOLESKY), drifts apart exponentially in step 2 (RIPOBA at small damping is chaotic on these
scenes, none of which converges within the 50 iterations), takes a different accept/reject decision somewhere
between trial 38 and 90, and ends 1-5 % away on trafalgar-257.  So:
  (a) step 1: every trial within max(1e-9, 5 x the reference's own running deviation, 5 x the reference's own worst
      step-1 deviation), identical accept/reject decisions and linear-solver iteration counts (the last term only
      matters for PCG / CHOLESKY, where the reference reproduces itself to 3e-8 / 5e-8 and our dense-S assembly
      uses atomics like the reference's own scatter-adds);
  (b) step 2 up to the first trial where either run (ours, or the reference's 8-thread one) decides
      differently from the 1-thread reference: within max(1e-9, 50 x the reference's running deviation);
  (c) final cost: within max(1e-6, 50 x the reference's final deviation) if no run took a different decision,
      otherwise within 5 % (different trajectories of a non-converged chaotic iteration).
DESIGN.md 5 has the measured numbers."""
import pytest

import povar_testlib as common
from povar_b200 import capi

pytestmark = pytest.mark.gpu


def _run(name):
    meta = common.traces_large()["traces"][name]
    kw = common.flags_to_options(meta["flags"])
    hp = capi.HostProblem.read(common.golden_file(meta["shape"]))
    s = capi.Solver(hp, capi.default_options(verbosity_level=0, **kw))
    its, summary = s.bundle_adjust()
    s.close()
    return meta, its, summary


def compare_with_reference_runs(name, meta, cost, succ, lin):
    ref, ref8 = meta["threads1"], meta["threads8"]
    k2 = common.step2_start(ref["iteration"])
    n = min(len(cost), len(ref["cost"]), len(ref8["cost"]))
    assert n > k2, f"{name}: the run ended inside step 1"

    def dev(c, i):
        return abs(c[i] - ref["cost"][i]) / abs(ref["cost"][i])

    drift = 0.0
    worst1 = 0.0
    step1_ref = max(dev(ref8["cost"], i) for i in range(k2))
    for i in range(k2):                                                     # (a)
        drift = max(drift, dev(ref8["cost"], i))
        d = dev(cost, i)
        worst1 = max(worst1, d)
        assert d <= max(1e-9, 5.0 * drift, 5.0 * step1_ref), \
            f"{name} step-1 trial {i}: rel {d:.2e} (reference's own drift {drift:.1e}, over step 1 {step1_ref:.1e})"
        assert bool(succ[i]) == bool(ref["step_is_successful"][i]), f"{name} step-1 trial {i}: accept/reject differs"
        assert int(lin[i]) == int(ref["linear_solver_iterations"][i]), f"{name} step-1 trial {i}: linear iterations"
    flip_ours = flip_ref8 = None
    worst_ratio = 0.0
    for i in range(k2, n):                                                  # (b)
        if bool(succ[i]) != bool(ref["step_is_successful"][i]):
            flip_ours = i
        if bool(ref8["step_is_successful"][i]) != bool(ref["step_is_successful"][i]):
            flip_ref8 = i
        if flip_ours is not None or flip_ref8 is not None:
            break
        drift = max(drift, dev(ref8["cost"], i))
        d = dev(cost, i)
        worst_ratio = max(worst_ratio, d / max(drift, 1e-12))
        assert d <= max(1e-9, 50.0 * drift), f"{name} step-2 trial {i}: rel {d:.2e} (reference's own drift {drift:.1e})"
    d_final = abs(cost[-1] - ref["cost"][-1]) / abs(ref["cost"][-1])       # (c)
    d8_final = abs(ref8["cost"][-1] - ref["cost"][-1]) / abs(ref["cost"][-1])
    if flip_ours is None and flip_ref8 is None:
        assert len(cost) == len(ref["cost"])
        assert d_final <= max(1e-6, 50.0 * d8_final), f"{name}: final cost rel {d_final:.2e}"
    else:
        assert d_final <= 5e-2, f"{name}: final cost rel {d_final:.2e}"
    return {"step1_worst": worst1, "first_flip_ours": flip_ours, "first_flip_ref8": flip_ref8,
            "worst_vs_ref_drift": worst_ratio, "final": d_final, "final_ref8": d8_final}


@pytest.mark.parametrize("name", ["trafalgar257_povar", "trafalgar257_poba", "trafalgar257_pcg",
                                  "trafalgar257_cholesky", "venice89_poba", "venice1778_povar_cauchy"])
def test_baseline_config_trace_matches_reference(name):
    if name not in common.traces_large()["traces"]:
        pytest.skip("trace not generated (tools/make_golden_large.py)")
    meta, its, summary = _run(name)
    stats = compare_with_reference_runs(name, meta, [e.cost for e in its], [e.step_is_successful for e in its],
                                        [e.linear_solver_iterations for e in its])
    print(f"{name}: {len(its)} trials; " + ", ".join(f"{k}={v:.2e}" if isinstance(v, float) else f"{k}={v}"
                                                       for k, v in stats.items()))
