"""Parity with the reference at BASELINE.json's sizes: full two-step traces of the CUDA path against
`bal_ref --num-threads 1` on the same (regenerated, hash-checked) data_custom files.

  c2  trafalgar-257 shape, all four step-1 solvers    c3  venice-89 shape, POWER_SCHUR_COMPLEMENT
  c4  venice-1778 shape, CAUCHY (the benchmark configuration)

Bars (povar_testlib.assert_trace_close): every step-1 trial and the first step-2 trials within 1e-9,
later ones within 1e-6, same accept/reject decisions and linear-solver iteration counts -- relaxed only
from the trial on which the reference's own 8-thread run has left its 1-thread run (stored next to it;
at these sizes its RIPOBA tail is chaotic: the two reference runs end 1-5 % apart on trafalgar-257)."""
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


@pytest.mark.parametrize("name", ["trafalgar257_povar", "trafalgar257_poba", "trafalgar257_pcg",
                                  "trafalgar257_cholesky", "venice89_poba", "venice1778_povar_cauchy"])
def test_baseline_config_trace_matches_reference(name):
    if name not in common.traces_large()["traces"]:
        pytest.skip("trace not generated (tools/make_golden_large.py)")
    meta, its, summary = _run(name)
    ref = meta["threads1"]
    k2 = common.step2_start(ref["iteration"])
    worst = common.assert_trace_close(meta, [e.cost for e in its], [e.step_is_successful for e in its],
                                      [e.linear_solver_iterations for e in its], label=name)
    # step 1 is stable in the reference at every size: per-trial cost within 1e-9, same decisions
    assert len(its) >= k2
    for i in range(k2):
        assert abs(its[i].cost - ref["cost"][i]) <= 1e-9 * ref["cost"][i], (i, its[i].cost, ref["cost"][i])
        assert bool(its[i].step_is_successful) == bool(ref["step_is_successful"][i])
        assert its[i].linear_solver_iterations == ref["linear_solver_iterations"][i]
    # (the first trials of step 2 are held to 1e-9 by assert_trace_close unless the reference's own two runs
    #  have already separated there)
    print(f"{name}: {len(its)} trials (reference {len(ref['cost'])}), worst deviation inside the bars {worst:.2e}, "
          f"final {its[-1].cost:.9e} vs {ref['cost'][-1]:.9e}")
