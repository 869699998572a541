"""world_size-2 check of the landmark sharding on CPU (gloo): the per-shard quantities the GPU
path all-reduces (cost sums, Jp^T Jp Kronecker sums, the camera-sized vector of an E0 product)
add up to the unsharded ones.  Shards come from the C ABI's povar_partition_landmarks; the
arithmetic is the oracle's (this is a test of the host-side plan, not of the kernels)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import povar_testlib as common
from oracle import povar_oracle as O
from povar_b200 import capi


def _shard_problem(hp, rank, world, P, X):
    sh = hp.shard(rank, world)
    lm = np.repeat(np.arange(sh.num_lms), np.diff(sh.lm_ptr)).astype(np.int32)
    prob = O.Problem(P=P.copy(), X=X[sh.lm_begin:sh.lm_end].copy(), Xh=np.zeros((sh.num_lms, 4)),
                     lm_ptr=sh.lm_ptr.copy(), obs_cam=sh.obs_cam.copy(), obs_lm=lm, uv=sh.obs_uv.copy())
    return prob


def _worker(rank, world, port, path, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    hp = capi.HostProblem.read(path)
    full = O.load_bal(path)
    opt = O.Options(alpha=0.1, power_sc_iterations=20)
    O.init_varproj(full, opt.alpha)
    prob = _shard_problem(hp, rank, world, full.P, full.X)
    # cost: 3 scalars
    ri = O.cost_pose(prob, opt)
    t = torch.tensor([ri.err_all, ri.rsum_all, float(ri.n_all)], dtype=torch.float64)
    dist.all_reduce(t)
    # linearisation: per-camera Jp^T Jp diagonal (what the pose scales are made of)
    r, Jp, Jl = O.pose_jacobians(prob, opt.alpha)
    diag2 = np.zeros((prob.C, 12))
    O.cam_scatter(diag2, prob.obs_cam, np.sum(Jp * Jp, axis=1))
    d = torch.from_numpy(diag2)
    dist.all_reduce(d)
    # one E0 product with local Hll^-1 (landmarks are whole inside a shard)
    Hll = O.seg_sum(np.einsum("nri,nrj->nij", Jl, Jl), prob.lm_ptr)
    x = np.random.default_rng(3).normal(size=(prob.C, 12))
    e0 = torch.from_numpy(O.right_mul_e0(Jp, Jl, O.inv3_cofactor(Hll), prob, x))
    dist.all_reduce(e0)
    if rank == 0:
        np.savez(os.path.join(out_dir, "reduced.npz"), cost=t.numpy(), diag2=d.numpy(), e0=e0.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sums_equal_unsharded(tmp_path):
    path = common.golden_file("small")
    world = 2
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, path, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "reduced.npz")
    full = O.load_bal(path)
    opt = O.Options(alpha=0.1, power_sc_iterations=20)
    O.init_varproj(full, opt.alpha)
    ri = O.cost_pose(full, opt)
    assert got["cost"][2] == full.nnz
    assert abs(got["cost"][0] - ri.err_all) <= 1e-13 * ri.err_all
    assert abs(got["cost"][1] - ri.rsum_all) <= 1e-13 * ri.rsum_all
    r, Jp, Jl = O.pose_jacobians(full, opt.alpha)
    diag2 = np.zeros((full.C, 12))
    O.cam_scatter(diag2, full.obs_cam, np.sum(Jp * Jp, axis=1))
    assert common.rel(got["diag2"], diag2) < 1e-13
    Hll = O.seg_sum(np.einsum("nri,nrj->nij", Jl, Jl), full.lm_ptr)
    x = np.random.default_rng(3).normal(size=(full.C, 12))
    e0 = O.right_mul_e0(Jp, Jl, O.inv3_cofactor(Hll), full, x)
    assert common.rel(got["e0"], e0) < 1e-12
