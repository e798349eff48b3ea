"""world_size-2 gloo test of the N>1 host path (sharding + the single all-gather), CPU only.
The solve itself is injected (the CPU oracle), so what is checked is the distributed plumbing:
sharded result == unsharded result, including a ragged split."""
import os
import socket

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from helpers import O, scene_case, oracle_level_inputs, rel_fro
from banet_b200 import dist as bdist


def test_shard_range_covers_everything():
    for nb in (1, 5, 32, 257):
        for world in (1, 2, 3, 8):
            spans = [bdist.shard_range(nb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    R = torch.randn(4, 3, 3); T = torch.randn(4, 3, 1); W = torch.randn(4, 7, 1)
    r, t, w = bdist.unpack_solution(bdist.pack_solution(R, T, W), 7)
    assert torch.equal(r, R) and torch.equal(t, T) and torch.equal(w, W)
    r, t, w = bdist.unpack_solution(bdist.pack_solution(R, T, None), 0)
    assert w is None and torch.equal(r, R)


def _solve(conv1, conv2, fx, fy, ox, oy, p, D, B, R, T, W):
    opts = O.IterOptions(lambda_override=torch.full((conv1.shape[0],), 0.05, dtype=torch.float64))
    for _ in range(2):
        R, T, W = O.bundle_iteration(conv1, conv2, fx, fy, ox, oy, p, D, B, R, T, W, None, opts)
    return R, T, W


def _worker(rank, world, port, nb, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    sc = scene_case(nb=nb, C=4, K=3, level_ids=(3,), seed=31, H=24, W=32)
    a = oracle_level_inputs(sc.levels[0])

    def shard(lo, hi):
        return tuple(a[k][lo:hi] for k in ("conv1", "conv2", "fx", "fy", "ox", "oy", "p", "D", "B")) + (sc.R0[lo:hi], sc.T0[lo:hi], sc.W0[lo:hi])

    R, T, W = bdist.solve_sharded(_solve, shard, nb)
    if rank == 0:
        q.put((R, T, W))
    td.destroy_process_group()


@pytest.mark.parametrize("nb", [4, 5])
def test_sharded_solve_equals_single_process(nb):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nb, q)) for r in range(2)]
    for p in procs:
        p.start()
    R, T, W = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc = scene_case(nb=nb, C=4, K=3, level_ids=(3,), seed=31, H=24, W=32)
    a = oracle_level_inputs(sc.levels[0])
    oR, oT, oW = _solve(*[a[k] for k in ("conv1", "conv2", "fx", "fy", "ox", "oy", "p", "D", "B")], sc.R0, sc.T0, sc.W0)
    assert rel_fro(R, oR) < 1e-12 and rel_fro(T, oT) < 1e-12 and rel_fro(W, oW) < 1e-12
