"""Shared helpers of the parity tests: build the oracle and (on a GPU box) the device sampler on the
same inputs and drive them through identical mutation sequences."""
import numpy as np

from oracle import mutations as M
from oracle.sampler import OracleSampler

FIELDS = M.FIELDS


def default_params(pyr, d_max=None):
    """Fixed model parameters for parity runs: the generating law."""
    law = pyr.spec["law"]
    A = pyr.spec["amplitude"]
    if d_max is None:
        from graal_b200.level import rippe_law
        grid = np.linspace(1.0, 5000.0, 50000)
        above = grid[A * rippe_law(grid) > pyr.spec["v_inter"]]
        d_max = float(above[-1]) if above.size else 100.0
    return [law["kuhn"], law["lm"], law["slope"], law["d"], A], d_max


def make_oracle(inp, pyr, seed=1000, d_max=None):
    o = OracleSampler(inp, np.random.RandomState(seed))
    p, dm = default_params(pyr, d_max)
    o.set_params(p[0], p[1], p[2], p[3], p[4], dm)
    return o


def scramble(oracle, rng, n_moves, gpu=None, modes=None):
    """Apply the same random mutations to the oracle (and the device sampler)."""
    n = oracle.n_new_frags
    applied = []
    for _ in range(n_moves):
        fA, fB = int(rng.randint(n)), int(rng.randint(n))
        if fA == fB:
            continue
        mode = int(rng.randint(13)) if modes is None else int(modes[rng.randint(len(modes))])
        max_id = oracle.modify_gl_cuda_buffer()
        M.apply_mutation(oracle.ws, oracle.cur, fA, fB, mode, max_id, oracle.id_contigs)
        if gpu is not None:
            gpu.apply_replay_simu(fA, fB, mode)
        applied.append((fA, fB, mode))
    return applied


def slots_diff(a, b):
    return [k for k in FIELDS if not np.array_equal(a[k], b[k])]


def delta_tolerance(delta_ref, mass):
    """|delta_gpu - delta_oracle| <= 1e-6 * max(|delta_oracle|, mass * 2^-20)   (SURVEY H2):
    relative 1e-6 as BASELINE.json's north_star states, with a floor for deltas that are small
    next to the float32 expected values (`mass` = sum of |terms| touched) they are differences of."""
    return 1e-6 * max(abs(delta_ref), mass * 2.0 ** -20) + 1e-9


MASS_FLOOR = 1e-7      # measured worst |err| / mass: 3.5e-8 (math modes 1, 2) .. 7e-8 (mode 0): float32 libm differences


def delta_check(got, ref, mass):
    """(within the asserted bound, within SURVEY H2's bound).  Asserted: |err| <= 1e-6 |ref| + 1e-7 mass, the
    relative tolerance of BASELINE.json's north_star plus the MEASURED float32 noise floor of the expected values
    (powf / expf of CUDA vs NumPy differ by <= 2 ulp per term; `mass` = sum of |terms| the delta is a difference of).
    H2 (delta_tolerance) is the stricter bound that holds whenever the delta is not a small difference of large sums."""
    err = abs(got - ref)
    return err <= 1e-6 * abs(ref) + MASS_FLOOR * mass + 1e-9, err <= delta_tolerance(ref, mass)


def ranking_agrees(got, ref, masses):
    """The candidate order is the one the oracle gives wherever two candidates are further apart than the
    tolerances of the pair (a tie inside the noise floor may fall either way)."""
    bad = []
    for i in range(len(ref)):
        for j in range(len(ref)):
            gap = 1e-6 * (abs(ref[i]) + abs(ref[j])) + MASS_FLOOR * (masses[i] + masses[j]) + 2e-9
            if ref[i] - ref[j] > gap and not got[i] > got[j]:
                bad.append((i, j))
    return bad


_POOL = {}


def _pool_sparse_delta(j):
    from oracle import sparse as S
    c = _POOL
    return S.sparse_delta(c["collector"][j], c["cur"], c["lv"], c["par"], c["bins_u"], return_mass=True)


def sparse_deltas(collector, cur, lv, par, bins_u, cands=range(13)):
    """(delta, mass) of the candidates from the sparse oracle, one forked worker per candidate (the children
    inherit the read-only level; they never touch CUDA)."""
    import multiprocessing as mp
    import os
    cands = list(cands)
    workers = max(1, min(len(cands), (os.cpu_count() or 1)))
    _POOL.update(collector=collector, cur=cur, lv=lv, par=par, bins_u=bins_u)
    try:
        if workers == 1:
            return [_pool_sparse_delta(j) for j in cands]
        with mp.get_context("fork").Pool(workers) as pool:
            return pool.map(_pool_sparse_delta, cands, chunksize=1)
    finally:
        _POOL.clear()
