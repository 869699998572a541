"""Landmark sharding over 2 GPUs (one process per GPU, NCCL all-reduce of the camera-sized vectors):
the sharded solve must reproduce the single-GPU trace up to summation order.  Needs 2 GPUs
(`gpurun --gpus 2`); skipped otherwise."""
import os
import sys

import numpy as np
import pytest
import torch

import povar_testlib as common

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, path, kw, out_dir, exchange):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["POVAR_PEER_EXCHANGE"] = exchange     # "1": peer memory or fail; "0": ncclAllReduce
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from povar_b200 import capi
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ids = [capi.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    hp = capi.HostProblem.read(path).shard(rank, world)
    s = capi.Solver(hp, capi.default_options(verbosity_level=0, **kw), capi.make_comm(rank, world, rank, ids[0]))
    assert s.peer_exchange_active() == (exchange == "1")
    its, summary = s.bundle_adjust()
    P, X = s.get_state(capi.STATE_JOINT)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), cost=[e.cost for e in its],
             succ=[e.step_is_successful for e in its], lin=[e.linear_solver_iterations for e in its], P=P, X=X,
             lm_begin=hp.lm_begin, lm_end=hp.lm_end)
    s.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,exchange", [("small_povar", "1"), ("small_povar", "0"), ("ladybug49_poba", "1"),
                                           ("ladybug49_pcg_ripcg", "1")])
def test_two_gpu_shards_reproduce_single_gpu_trace(name, exchange, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from povar_b200 import capi
    meta = common.traces()["traces"][name]
    kw = common.flags_to_options(meta["flags"])
    path = common.golden_file(meta["shape"])
    world = 2
    mp.spawn(_worker, args=(world, 29700 + os.getpid() % 200, path, kw, str(tmp_path), exchange), nprocs=world,
             join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    # replicated control flow: both ranks log the same trace, bit for bit
    assert np.array_equal(r0["cost"], r1["cost"]) and np.array_equal(r0["lin"], r1["lin"])
    assert np.array_equal(r0["P"], r1["P"])
    hp = capi.HostProblem.read(path)
    s = capi.Solver(hp, capi.default_options(verbosity_level=0, **kw))
    its, _ = s.bundle_adjust()
    P, X = s.get_state(capi.STATE_JOINT)
    s.close()
    common.assert_trace_close(meta, list(r0["cost"]), list(r0["succ"]), list(r0["lin"]), label=name + " x2")
    k2 = common.step2_start(meta["threads1"]["iteration"])
    n = min(len(its), len(r0["cost"]))
    for i in range(min(n, k2 + 6)):
        assert abs(its[i].cost - r0["cost"][i]) <= 1e-9 * abs(its[i].cost)
    assert r0["lm_end"] == r1["lm_begin"] and r1["lm_end"] == hp.num_lms


def test_bal_binary_with_two_gpus(tmp_path):
    """`bal --num-gpus 2` (forked ranks, NCCL id through pipes, peer exchange between the two processes) writes the
    same trace as the single-GPU binary, up to summation order."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import json
    import subprocess
    from povar_b200 import build
    meta = common.traces()["traces"]["small_povar"]
    logs = []
    for n in (1, 2):
        log = tmp_path / f"ba_log_{n}.json"
        res = subprocess.run([build.BAL, "--input", common.golden_file("small"), "--alpha", "0.1",
                              "--power-sc-iterations", "20", "--num-gpus", str(n), "--verbosity-level", "0",
                              "--log-log-path", str(log)], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        logs.append(json.loads(log.read_text()))
    common.assert_trace_close(meta, logs[1]["cost"], logs[1]["step_is_successful"],
                              logs[1]["linear_solver_iterations"], label="bal x2")
    k2 = common.step2_start(meta["threads1"]["iteration"])
    for i in range(min(k2 + 6, len(logs[0]["cost"]), len(logs[1]["cost"]))):
        assert abs(logs[0]["cost"][i] - logs[1]["cost"][i]) <= 1e-9 * abs(logs[0]["cost"][i])
