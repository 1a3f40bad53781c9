"""Replica exchange (the only multi-GPU step of the path): world_size-2 gloo on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graal_b200 import replica as R


def test_swap_decisions_are_deterministic_and_valid():
    T = R.temperature_ladder(8)
    assert T[0] == 1.0 and abs(T[3] - 1.25 ** 3) < 1e-12
    idx = np.arange(8)
    L = np.array([-100.0, -50.0, -300.0, -10.0, -20.0, -500.0, -5.0, -7.0])
    a, log_a = R.swap_decisions(L, idx, T, 0, 42)
    b, log_b = R.swap_decisions(L, idx, T, 0, 42)
    assert np.array_equal(a, b) and log_a == log_b
    assert sorted(a.tolist()) == list(range(8))                 # labels stay a permutation
    # even round proposes (0,1),(2,3),(4,5),(6,7); a better likelihood at the hotter label always moves down
    assert [k for k, *_ in log_a] == [0, 2, 4, 6]
    assert log_a[0][1] and a[1] == 0 and a[0] == 1             # L[1] > L[0]: accepted for sure
    c, log_c = R.swap_decisions(L, a, T, 1, 42)
    assert [k for k, *_ in log_c] == [1, 3, 5]
    # non-finite likelihoods never swap
    Lnan = L.copy(); Lnan[0] = np.nan
    d, log_d = R.swap_decisions(Lnan, idx, T, 0, 42)
    assert not log_d[0][1]


def test_acceptance_rate_matches_metropolis():
    T = np.array([1.0, 2.0])
    acc = 0
    for r in range(0, 4000, 2):
        _, log = R.swap_decisions([0.0, -1.0], [0, 1], T, r, 9)
        acc += log[0][1]
    expect = np.exp((1.0 - 0.5) * (-1.0))
    assert abs(acc / 2000.0 - expect) < 0.03


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local = 2
    rex = R.ReplicaExchange(n_local, R.temperature_ladder(n_local * world), exchange_every=5, seed=123)
    rng = np.random.RandomState(1000 + rank)
    temps = []
    for step in range(1, 41):
        logliks = -100.0 * rng.rand(n_local) - 10.0 * np.array([rex.temperature(c) for c in range(n_local)])
        rex.maybe_exchange(step, logliks)
        temps.append([rex.temperature(c) for c in range(n_local)])
    rex.consume()                                    # the round posted at the last boundary
    temps[-1] = [rex.temperature(c) for c in range(n_local)]
    q.put((rank, rex.temp_index.tolist(), rex.n_collectives, rex.history, temps[-1]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_exchange_is_consistent():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, idx0, n0, h0, t0), (r1, idx1, n1, h1, t1) = res
    assert idx0 == idx1 and sorted(idx0) == [0, 1, 2, 3]        # every rank holds the same permutation
    assert n0 == n1 == 8                                        # one all_gather per exchange round
    assert h0 == h1
    assert any(acc for rnd in h0 for (_, acc, _, _) in rnd)
    T = R.temperature_ladder(4)
    assert t0 == [T[idx0[0]], T[idx0[1]]] and t1 == [T[idx1[2]], T[idx1[3]]]


def test_single_process_fallback_and_attach():
    rex = R.ReplicaExchange(3, R.temperature_ladder(3), exchange_every=2, seed=1)

    class Dummy:
        def temperature(self, t=0, n=1):
            return 1.0
    s = R.attach(Dummy(), rex, 2)
    assert s.temperature() == 1.25 ** 2
    rex.exchange([-5.0, -1.0, -3.0])
    assert sorted(rex.temp_index.tolist()) == [0, 1, 2] and s.temperature() == R.temperature_ladder(3)[rex.temp_index[2]]
    with pytest.raises(ValueError):
        R.ReplicaExchange(2, R.temperature_ladder(3))


def test_exchange_is_applied_one_boundary_later():
    """post / consume: the table gathered at a boundary decides the labels at the NEXT step boundary (nothing blocks)."""
    rex = R.ReplicaExchange(2, R.temperature_ladder(2), exchange_every=3, seed=5)
    L = [-100.0, -1.0]                                # the hotter chain is far better: the swap is certain
    for step in (1, 2):
        assert rex.maybe_exchange(step, L) is None
    assert rex.maybe_exchange(3, L) is None           # posted, not applied yet
    assert rex.temp_index.tolist() == [0, 1]
    log = rex.maybe_exchange(4, L)                    # applied at the next boundary
    assert log is not None and log[0][1] and rex.temp_index.tolist() == [1, 0]
    st = rex.stats()
    assert st["posted"] == 1 and st["blocking"] is False
    rex.warm_up()                                     # changes nothing
    assert rex.temp_index.tolist() == [1, 0] and rex.round_id == 1 and rex.stats()["posted"] == 1
    rex.reset()
    assert rex.temp_index.tolist() == [0, 1] and rex.round_id == 0
