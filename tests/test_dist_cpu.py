"""CPU, world_size 2 over gloo: the host-side sharding logic of the multi-GPU path.
The data path has no collective -- instances are sharded along B -- so what is checked is (i) the shard bounds tile
[0,B) exactly, (ii) slicing ONE global uniform vector per rank reproduces the single-rank systematic ancestors, and
(iii) the metric all-reduce returns the global per-sweep means on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from apgs_b200 import dist as adist
from oracle import resampler as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_tile_the_batch():
    for B in (1, 7, 20, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [adist.shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (a0, a1), (b0, b1) in zip(spans[:-1], spans[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, S, B, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                       # every rank builds the same global problem, then takes its shard
        log_w = torch.randn(S, B) * 2
        u = torch.rand(B)
        dens = torch.randn(3, S, B)
        lw_r = adist.shard_instances(log_w, rank, world, dim=1).contiguous()
        u_r = adist.shard_uniforms(u, rank, world)
        w_r = torch.softmax(lw_r, 0)
        anc_r = orc.systematic_ancestors(w_r.numpy(), u_r.numpy())
        ess_r = (1. / (w_r ** 2).sum(0)).unsqueeze(0).repeat(3, 1)
        logz_r = (torch.logsumexp(lw_r, 0) - np.log(S)).unsqueeze(0).repeat(3, 1)
        trace = {"ess": ess_r, "density": adist.shard_instances(dens, rank, world, dim=2), "log_z": logz_r}
        m = adist.allreduce_metrics(trace)
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), anc=anc_r, ess=m["ess"].numpy(), density=m["density"].numpy(),
                 count=m["count"].numpy(), log_z=m["log_z"].numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_run_equals_single_rank(tmp_path):
    S, B, world = 16, 11, 2
    mp.spawn(_worker, args=(world, _free_port(), S, B, str(tmp_path)), nprocs=world, join=True)
    torch.manual_seed(0)
    log_w = torch.randn(S, B) * 2
    u = torch.rand(B)
    dens = torch.randn(3, S, B)
    w = torch.softmax(log_w, 0)
    anc = orc.systematic_ancestors(w.numpy(), u.numpy())
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    assert np.array_equal(np.concatenate([p["anc"] for p in parts], 1), anc)
    ess = (1. / (w ** 2).sum(0)).mean().item()
    log_z = (torch.logsumexp(log_w, 0) - np.log(S)).double().mean().item()
    for p in parts:
        np.testing.assert_allclose(p["log_z"], log_z, rtol=1e-6)
        np.testing.assert_allclose(p["ess"], ess, rtol=1e-6)
        np.testing.assert_allclose(p["density"], dens.double().mean(1).mean(-1).numpy(), rtol=1e-6, atol=1e-9)
        assert (p["count"] == B).all()
