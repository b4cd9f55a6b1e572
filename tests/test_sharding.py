"""Multi-GPU host logic on CPU: world_size-2 `gloo` processes shard a batch, each evaluates its shard with the
oracle's marginalised likelihood as the stand-in evaluator, and the gathered result must equal the single-process
evaluation (the path has no data-path collective; only the per-point results are gathered)."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_bounds_cover_the_batch():
    from eftpipe_b200.shard import shard_bounds

    for n in (0, 1, 7, 64, 1000, 65536):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _points(n):
    rng = np.random.default_rng(1234)
    ndata, ng = 12, 3
    A = rng.normal(size=(ndata, ndata))
    invcov = np.linalg.inv(A @ A.T + ndata * np.eye(ndata))
    data = rng.normal(size=ndata)
    PNG = rng.normal(size=(n, ndata))
    PG = rng.normal(size=(n, ng, ndata))
    return PNG, PG, data, invcov


def _local_eval(lo, hi, n):
    import torch

    sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
    import pybird_oracle as orc

    PNG, PG, data, invcov = _points(n)
    logp = [orc.marginalized_logp(PNG[i], PG[i], data, invcov) for i in range(lo, hi)]
    status = [i % 5 for i in range(lo, hi)]
    return torch.tensor(logp, dtype=torch.float64), torch.tensor(status, dtype=torch.int32)


def _worker(rank, world, port, n, q):
    import torch.distributed as dist

    from eftpipe_b200.shard import evaluate_sharded, shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        logp, status = evaluate_sharded(lambda lo, hi: _local_eval(lo, hi, n), n)
        q.put((rank, shard_bounds(n, world, rank), logp.numpy(), status.numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [9, 16])
def test_two_rank_gloo_gather_equals_single_process(n):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_logp, ref_status = _local_eval(0, n, n)
    bounds = sorted(g[1] for g in got)
    assert bounds[0][0] == 0 and bounds[0][1] == bounds[1][0] and bounds[1][1] == n
    for _, _, logp, status in got:  # every rank holds the full gathered vectors
        np.testing.assert_array_equal(logp, ref_logp.numpy())
        np.testing.assert_array_equal(status, ref_status.numpy())
