"""The reference's own driver on the GPU library: oracle/_ref/bal_ref_b200 is the reference program built from its
unmodified sources with integration/linearizor_factory_b200.cpp in place of its factory TU, i.e. the reference's
`bundle_adjust_manual` (both LM loops, backup / restore, step-2 normalisation on the host BalProblem) calling
libpovar_b200.so through its `Linearizor` interface.  Its trace must match the reference's own.

Built and linked in the authoring container (`make -C oracle plugin`, no GPU there)."""
import json
import os
import subprocess

import pytest

import povar_testlib as common

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "oracle", "_ref", "bal_ref_b200")


@pytest.mark.parametrize("name", ["tiny_povar", "small_povar", "small_poba"])
def test_reference_driver_runs_on_the_gpu_library(name, tmp_path):
    if not os.path.exists(PLUGIN):
        pytest.skip("oracle/_ref/bal_ref_b200 not built (make -C oracle plugin)")
    meta = common.traces()["traces"][name]
    merged = {"--alpha": "0.1", "--power-sc-iterations": "20"}
    merged.update(dict(zip(meta["flags"][0::2], meta["flags"][1::2])))
    cmd = [PLUGIN, "--input", common.golden_file(meta["shape"]), "--num-threads", "1"]
    for k, v in merged.items():
        cmd += [k, v]
    log = tmp_path / "ba_log.json"
    cmd += ["--log-log-path", str(log)]
    res = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    data = json.loads(log.read_text())
    common.assert_trace_close(meta, data["cost"], data["step_is_successful"], data["linear_solver_iterations"],
                              label=name + " (reference driver + LinearizorB200)")
