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


def test_reference_driver_logs_the_library_phase_times(tmp_path):
    """LinearizorB200 copies the library's CUDA-event phase times (povar_get_timings) into the IterationSummary
    fields the reference's own linearizors fill (solver/linearizor_power_varproj.cpp:61-306): the reference's
    ba_log.json then has its usual time columns."""
    if not os.path.exists(PLUGIN):
        pytest.skip("oracle/_ref/bal_ref_b200 not built (make -C oracle plugin)")
    log = tmp_path / "ba_log.json"
    res = subprocess.run([PLUGIN, "--input", common.golden_file("small"), "--num-threads", "1", "--alpha", "0.1",
                          "--power-sc-iterations", "20", "--max-num-iterations-step-1", "5",
                          "--max-num-iterations-step-2", "3", "--log-log-path", str(log)],
                         cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    data = json.loads(log.read_text())
    for col in ("jacobian_evaluation_time", "prepare_time", "solve_reduced_system_time", "back_substitution_time"):
        assert col in data, col
        vals = [v for v, it in zip(data[col], data["iteration"]) if it > 0]
        assert vals and all(v >= 0 for v in vals) and sum(vals) > 0, (col, vals)


@pytest.mark.parametrize("name", ["small_povar", "small_poba"])
def test_reference_driver_on_two_shards(name, tmp_path):
    """Two processes of the reference's own driver, each holding one landmark shard on the GPU library
    (POVAR_PLUGIN_WORLD / RANK / DEVICE / ID_FILE, integration/linearizor_b200.hpp); both ranks share device 0
    through the host rendezvous, so this runs on a one-GPU box.  Both must log the reference's golden trace."""
    if not os.path.exists(PLUGIN):
        pytest.skip("oracle/_ref/bal_ref_b200 not built (make -C oracle plugin)")
    meta = common.traces()["traces"][name]
    merged = {"--alpha": "0.1", "--power-sc-iterations": "20"}
    merged.update(dict(zip(meta["flags"][0::2], meta["flags"][1::2])))
    procs, logs = [], []
    for rank in range(2):
        work = tmp_path / f"rank{rank}"
        work.mkdir()
        log = work / "ba_log.json"
        cmd = [PLUGIN, "--input", common.golden_file(meta["shape"]), "--num-threads", "1", "--log-log-path", str(log)]
        for k, v in merged.items():
            cmd += [k, v]
        env = dict(os.environ, POVAR_PLUGIN_WORLD="2", POVAR_PLUGIN_RANK=str(rank), POVAR_PLUGIN_DEVICE="0",
                   POVAR_PLUGIN_HOST_ID="1", POVAR_PLUGIN_ID_FILE=str(tmp_path / "comm.id"))
        procs.append(subprocess.Popen(cmd, cwd=work, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        logs.append(log)
    for p in procs:
        out, err = p.communicate(timeout=900)
        assert p.returncode == 0, err[-2000:]
    traces = [json.loads(l.read_text()) for l in logs]
    assert traces[0]["cost"] == traces[1]["cost"]          # replicated decisions, bit for bit
    common.assert_trace_close(meta, traces[0]["cost"], traces[0]["step_is_successful"],
                              traces[0]["linear_solver_iterations"], label=name + " (reference driver x2)")
