"""The host logic of a step against the reference's OWN Python lines (cuda_lib_gl.py), executed under Python 3
by tests/ref_host.py: the candidate draw of step_max_likelihood, return_neighbours, setup_distri_frags and
dist_inter_genome.  Same global NumPy RandomState on both sides: same draws, same results, same stream position
afterwards.  Runs where /root/reference exists."""
import types

import numpy as np
import pytest

import ref_host as RH
import helpers as H
from graal_b200 import sampler as GS
from graal_b200.level import prepare_sampler_inputs, build_synthetic_pyramid
from oracle import mutations as M
from oracle import sampler as OS

pytestmark = pytest.mark.skipif(not RH.available(), reason="reference sources not present")
NT = 13


def test_candidate_draw_matches_the_reference_lines():
    """cuda_lib_gl.py:1901-1941 (filter, threshold, temperature, np.random.choice) vs sampler._sample."""
    code = RH.block("step_max_likelihood", "scores_2_remove = []", "sample_out = np.random.choice(id_ok_4_sampling[0], 1, p=self.sub_score)[0]")
    gen = np.random.RandomState(0)
    n_draws = 0
    for case in range(400):
        n_nb = int(gen.randint(1, 5))
        score = -1e6 + gen.randn(NT * n_nb) * gen.choice([0.01, 1.0, 5.0, 40.0])
        if case % 7 == 0:
            score[gen.randint(score.size)] += 100.0                # one dominant candidate: argmax path
        if case % 11 == 0:
            score[:] = score[0]                                    # all equal
        F_t = float(gen.choice([1.0, 1.0, 1.7, 0.6]))
        seed = int(gen.randint(1 << 30))
        ref_self = types.SimpleNamespace(score=score.copy(), n_tmp_struct=NT, temperature=lambda t, n: F_t, sub_score=None)
        ns = {"self": ref_self, "np": np, "time": __import__("time"), "t": 0, "n_step": 1, "id_neighbours": list(range(n_nb))}
        np.random.seed(seed)
        exec(code, ns)
        ref_out, ref_next = int(ns["sample_out"]), np.random.rand()
        mine = types.SimpleNamespace(rng=np.random, _remove_cache={}, sub_score=None)
        np.random.seed(seed)
        got = GS.sampler._sample(mine, score.copy(), F_t)
        got_next = np.random.rand()
        assert got == ref_out and got_next == ref_next, (case, got, ref_out)
        # the oracle's restatement (what the golden trajectories were generated with) takes an explicit RandomState
        rs = np.random.RandomState(seed)
        o_out, _ = OS.sample_candidate(score.copy(), NT, rs, F_t)
        assert o_out == ref_out and rs.rand() == ref_next, (case, o_out, ref_out)
        n_draws += int(ref_next != np.random.RandomState(seed).rand())
    assert n_draws > 100                                           # the draw path (not only argmax) was exercised


def _mock_level(allow_repeats, blacklist):
    pyr = build_synthetic_pyramid([300_000, 200_000, 150_000, 90_000, 40_000, 6_000], 600, 3, seed=11, cis_rowsum=300.0, v_inter=0.05)
    kw = dict(allow_repeats=allow_repeats)
    if blacklist:
        kw["blacklist_contigs"] = (5,)
    return pyr, prepare_sampler_inputs(pyr, 2, **kw)


@pytest.mark.parametrize("allow_repeats,blacklist", [(False, False), (True, False), (False, True)])
def test_neighbour_tables_and_return_neighbours(allow_repeats, blacklist):
    """setup_distri_frags (:2363-2390) and return_neighbours (:2295-2331) as the reference wrote them."""
    pyr, inp = _mock_level(allow_repeats, blacklist)
    N, n = int(inp.n_frags), int(inp.n_new_frags)
    r, c, v = inp.level_coo
    dense = np.zeros((N, N), dtype=np.float32)
    dense[r, c] = v; dense[c, r] = v
    np.fill_diagonal(dense, 0)
    black_bins = sorted(set(int(inp.S_o_A_frags["id_d"][f]) for f in inp.id_frags_blacklisted))
    if black_bins:                                                 # the loader zeroes blacklisted rows / columns before the sampler sees them
        dense[black_bins, :] = 0; dense[:, black_bins] = 0
    ref_self = types.SimpleNamespace(n_frags=N, hic_matrix_sub_sampled=dense, n_neighbors=10)
    RH.method("setup_distri_frags")(ref_self)
    xk, pk = GS.neighbour_tables((r, c, v), N, black_bins, 10)
    for i in range(N):
        rp, rx = ref_self.distri_frags[i]["pk"], ref_self.distri_frags[i]["xk"]
        live = rp > 0
        assert np.array_equal(pk[i][: rp.size] > 0, live), i
        # the drawable columns: same probabilities (which column wins among EQUAL counts is the unstable argsort's choice)
        assert np.allclose(np.sort(pk[i][: rp.size][live]), np.sort(rp[live]), rtol=1e-6)
        top = np.sort(dense[i])[::-1][: int(live.sum()) + 1]
        if np.unique(top).size == top.size:                        # no ties down to the first excluded column: the order is defined
            assert np.array_equal(xk[i][: rp.size][live], rx[live]) and np.allclose(pk[i][: rp.size][live], rp[live], rtol=1e-6)
    # return_neighbours with the reference's tables on both sides (ties cannot matter then)
    disp = np.zeros(N, dtype=[("x", np.int32), ("y", np.int32)])
    d2 = np.asarray(inp.frag_dispatcher).reshape(-1, 2)
    disp["x"], disp["y"] = d2[:, 0], d2[:, 1]
    state = types.SimpleNamespace(id_d=np.asarray(inp.S_o_A_frags["id_d"]))
    ref_self = types.SimpleNamespace(gpu_vect_frags=state, n_neighbors=10, distri_frags=ref_self.distri_frags,
                                     id_frag_duplicated=list(np.asarray(inp.id_frag_duplicated)), frag_dispatcher=disp,
                                     collector_id_repeats=np.asarray(inp.collector_id_repeats), id_frags_blacklisted=list(inp.id_frags_blacklisted))
    ref_fn = RH.method("return_neighbours")
    ref_xk = np.stack([ref_self.distri_frags[i]["xk"] for i in range(N)])
    ref_pk = np.stack([ref_self.distri_frags[i]["pk"] for i in range(N)])
    dup = set(int(f) for f in np.asarray(inp.id_frag_duplicated))
    mine = types.SimpleNamespace(h_id_d=np.asarray(inp.S_o_A_frags["id_d"]), n_neighbors=10, distri_pk=ref_pk, distri_xk=ref_xk,
                                 _n_cand_cache={}, rng=np.random, _dup_set=dup, _black_set=set(int(f) for f in inp.id_frags_blacklisted),
                                 frag_dispatcher=d2, collector_id_repeats=np.asarray(inp.collector_id_repeats),
                                 _identity_dispatch=bool(len(dup) == 0 and np.all(d2[:, 1] - d2[:, 0] == 1) and
                                                         np.array_equal(np.asarray(inp.collector_id_repeats)[d2[:, 0]], np.arange(N))))
    gen = np.random.RandomState(3)
    for case in range(300):
        fA = int(gen.randint(n)); delta = int(gen.choice([1, 3, 5, 12])); seed = int(gen.randint(1 << 30))
        np.random.seed(seed)
        a = [int(e) for e in ref_fn(ref_self, fA, delta)]
        na = np.random.rand()
        np.random.seed(seed)
        b = GS.sampler.return_neighbours(mine, fA, delta)
        nb = np.random.rand()
        assert a == b and na == nb, (case, fA, delta, a, b)


def test_dist_inter_genome():
    """dist_inter_genome (:475-541): the reference loop vs the vectorised host function and the oracle's."""
    pyr, inp = _mock_level(True, False)
    o = H.make_oracle(inp, pyr)
    rng = np.random.RandomState(2)
    n = int(inp.n_new_frags)
    ref_fn = RH.method("dist_inter_genome")
    init = {k: np.asarray(inp.S_o_A_frags[k]).copy() for k in M.FIELDS}
    sub_id = np.asarray(inp.np_sub_frags_id).reshape(-1, 4)
    orientable = (sub_id[init["id_d"], 3] > 1).astype(np.int32)
    is_repeat = list(np.isin(init["id_d"], np.unique(init["id_d"][init["id"] != init["id_d"]])))
    n4 = len(np.unique(list(inp.id_frags_blacklisted) + list(np.nonzero(is_repeat)[0])))
    for rnd in range(4):
        H.scramble(o, rng, 25)
        cur = types.SimpleNamespace(copy_from_gpu=lambda: None, **{k: o.cur[k] for k in M.FIELDS})
        ref_self = types.SimpleNamespace(id_frags_blacklisted=list(inp.id_frags_blacklisted), n_new_frags=n, n_frags_4_dist=n4,
                                         is_repeat=is_repeat, np_init_prev=init["prev"], np_init_next=init["next"],
                                         np_init_ori=np.ones(n, dtype=np.int32), np_init_orientable=orientable)
        want = ref_fn(ref_self, cur)
        got = GS.dist_inter_genome(o.cur["prev"], o.cur["next"], o.cur["ori"], o.cur["id_d"], init["prev"], init["next"],
                                   np.ones(n, dtype=np.int32), orientable, list(inp.id_frags_blacklisted), np.asarray(is_repeat), n, n4)
        assert abs(got - want) < 1e-12, (rnd, got, want)
        assert abs(o.dist_inter_genome(o.cur) - want) < 1e-12
