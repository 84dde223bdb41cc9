"""Host logic of ping-sharded execution on CPU with the gloo backend (world_size 2 and 3): the global ping-bin
grid (pipeline.global_ping_edges) and the straddling-bin reduce (pipeline.straddle_reduce) give every rank the
same (sum, count) accumulators for the bins it touches as a single-process reduction over the whole volume."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_case(seed, P, bounds):
    """Deterministic 'volume': per ping a [C, nR, 4] contribution; ping times 1 s apart, 7 s bins."""
    rng = np.random.default_rng(seed)
    C, nR = 2, 5
    contrib = rng.integers(0, 1000, size=(P, C, nR, 4)).astype(np.float64)  # integers: sums are order independent
    t0 = np.datetime64("2018-07-01T00:00:03", "ns")
    pt = t0 + (np.arange(P) * 1_000_000_000).astype("timedelta64[ns]")
    return contrib, pt, bounds


def _worker(rank, world, port, seed, P, bounds, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from echopype_b200 import pipeline
        from echopype_b200.commongrid.utils import assign_bins

        contrib, pt, bounds = _global_case(seed, P, bounds)
        a, b = bounds[rank], bounds[rank + 1]
        edges = pipeline.global_ping_edges(pt[a:b], "7s", dist.group.WORLD)
        xb = assign_bins(pt[a:b], edges)
        lo, hi = int(xb.min()), int(xb.max())
        acc = torch.zeros((2, hi - lo + 1, 5, 4), dtype=torch.float64)
        for i, k in enumerate(xb):
            acc[:, k - lo] += torch.from_numpy(contrib[a + i])
        pipeline.straddle_reduce(acc, lo, hi, dist.group.WORLD)
        q.put((rank, lo, hi, acc.numpy(), edges.astype("datetime64[ns]").astype(np.int64)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,P,bounds", [
    (2, 40, [0, 20, 40]),       # boundary inside a 7 s bin
    (2, 42, [0, 21, 42]),
    (3, 30, [0, 10, 12, 30]),   # middle rank lies inside a single bin shared with both neighbours
    (3, 35, [0, 4, 18, 35]),
])
def test_straddle_reduce_matches_single_process(world, P, bounds):
    from echopype_b200.commongrid.utils import assign_bins, ping_time_edges

    seed = 7
    contrib, pt, _ = _global_case(seed, P, bounds)
    edges = ping_time_edges(pt, "7s")
    xb = assign_bins(pt, edges)
    want = np.zeros((2, len(edges) - 1, 5, 4))
    for i, k in enumerate(xb):
        want[:, k] += contrib[i]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, seed, P, bounds, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, lo, hi, acc, e in got:
        np.testing.assert_array_equal(e, edges.astype("datetime64[ns]").astype(np.int64))  # same global grid on every rank
        np.testing.assert_array_equal(acc, want[:, lo : hi + 1])


def test_straddle_neighbours_pairwise_and_fallback():
    from echopype_b200.pipeline import straddle_neighbours

    # four shards, every boundary inside a bin: each shared bin has exactly two holders
    w = [(0, 5), (5, 10), (10, 15), (15, 20)]
    assert straddle_neighbours(w, 0) == {"pairwise": True, "left": None, "right": 1}
    assert straddle_neighbours(w, 2) == {"pairwise": True, "left": 1, "right": 3}
    assert straddle_neighbours(w, 3) == {"pairwise": True, "left": 2, "right": None}
    # boundaries on bin edges: nothing is shared, still pairwise (no traffic at all)
    w = [(0, 4), (5, 9), (10, 14)]
    assert straddle_neighbours(w, 1) == {"pairwise": True, "left": None, "right": None}
    # a one-bin shard shared with one neighbour only
    w = [(0, 5), (5, 5), (6, 9)]
    assert straddle_neighbours(w, 1) == {"pairwise": True, "left": 0, "right": None}
    assert straddle_neighbours(w, 2) == {"pairwise": True, "left": None, "right": None}
    # the middle shard lies inside one bin shared with BOTH neighbours: three holders -> all-reduce path on every rank
    w = [(0, 5), (5, 5), (5, 9)]
    assert all(not straddle_neighbours(w, r)["pairwise"] for r in range(3))
    # an empty shard (-1, -1) is ignored
    w = [(0, 5), (-1, -1), (5, 9)]
    assert straddle_neighbours(w, 0) == {"pairwise": True, "left": None, "right": 2}
    assert straddle_neighbours(w, 2) == {"pairwise": True, "left": 0, "right": None}


def _plan_worker(rank, world, port, pings, ping_num, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from echopype_b200 import pipeline

        start = sum(pings[:rank])
        t0 = 1_530_000_000_000_000_000
        shard = (pings[rank], t0 + start * 10**9, t0 + (start + pings[rank] - 1) * 10**9, ping_num)
        try:
            plan = pipeline.straddle_plan(start // 20, (start + pings[rank] - 1) // 20, dist.group.WORLD, shard)
            q.put((rank, "ok", plan["pairwise"]))
        except ValueError as e:
            q.put((rank, "ValueError", str(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("pings,ping_num,ok", [([25, 35], 5, True), ([23, 37], 5, False), ([23, 37], 1, True)])
def test_sharded_plan_checks_ping_num_multiple(pings, ping_num, ok):
    """every rank but the last must hold a multiple of ping_num pings: otherwise ValueError on EVERY rank"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_plan_worker, args=(r, 2, port, pings, ping_num, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if ok:
        assert [g[1] for g in got] == ["ok", "ok"]
    else:
        assert [g[1] for g in got] == ["ValueError", "ValueError"]
        assert "multiple of ping_num" in got[0][2]
