"""The remaining entry points of the reference ``sampler`` surface (graal_b200/variants.py) on the device against the oracle's
restatement (oracle/variants.py, pinned to the reference's own lines by tests/test_reference_variants.py): per-mode builders,
per-neighbour scoring, the validation step that scores by FULL likelihoods, the older proposal rule and its step,
local_flip, the scramblers."""
import numpy as np
import pytest

from graal_b200.level import prepare_sampler_inputs
from oracle import mutations as M
import helpers as H

pytestmark = pytest.mark.gpu


def _pair(pyr, level, seed, **kw):
    from graal_b200.sampler import sampler
    inp = prepare_sampler_inputs(pyr, level, **kw)
    o = H.make_oracle(inp, pyr, seed=seed)
    g = sampler.from_inputs(inp, rng=np.random.RandomState(seed))
    p, dm = H.default_params(pyr)
    g.set_parameters(p, dm)
    return inp, o, g


def test_builders_and_per_neighbour_scores(small_pyramid):
    from graal_b200.sampler import CUR, CAND0
    inp, o, g = _pair(small_pyramid, 2, 31)
    rng = np.random.RandomState(8)
    H.scramble(o, rng, 30, g)
    max_id = int(o.modify_gl_cuda_buffer()); g.modify_gl_cuda_buffer()
    o.init_likelihood(); g.init_likelihood()
    n = o.n_new_frags
    for case in range(4):
        fA, fB = (int(x) for x in rng.choice(n, 2, replace=False))
        # one mode at a time, then the translocation family: the collector slots of the oracle
        for mode in range(9):
            M.pop_out_pop_in(o.ws, o.cur, fA, fB, mode, max_id)
            g.pop_out_pop_in(fA, fB, mode, max_id)
            assert H.slots_diff(o.ws.collector[mode], g.slot_to_host(CAND0 + mode)) == [], (fA, fB, mode)
        M.transloc(o.ws, o.cur, fA, fB, max_id)
        g.transloc(fA, fB, max_id)
        for mode in range(9, 13):
            assert H.slots_diff(o.ws.collector[mode], g.slot_to_host(CAND0 + mode)) == [], (fA, fB, mode)
        # stream_likelihood: score[13 x + j] = likelihood_t + delta_j
        o.score = np.zeros(26); o.delta = np.zeros(26)
        g.score = np.zeros(26)
        o.stream_likelihood(fA, fB, 1, o.likelihood_t, max_id)
        g.stream_likelihood(fA, None, None, fB, 1, g.likelihood_t, max_id)
        assert np.all(g.score[:13] == 0)
        assert np.allclose(o.score[13:], g.score[13:], rtol=1e-8), (fA, fB, np.abs(o.score - g.score).max())
        # the MH builders one family at a time == all_modifications_metropolis
        o.all_modifications_metropolis(fA, fB, max_id, True)
        for mode in range(6):
            g.pop_out_pop_in_4_mh(fA, fB, mode, max_id, True)
        g.split_4_mh(fA, max_id, True)
        g.paste_4_mh(fA, fB, max_id, True)
        g.transloc_4_mh(fA, fB, max_id, True)
        for j in range(13):
            assert H.slots_diff(o.ws.collector[j], g.slot_to_host(CAND0 + j)) == [], (fA, fB, j)
        so = o.compute_all_score_MH(fA, {fB}, True)
        sg = np.zeros(13)
        g.multi_likelihood_4_metropolis(fA, None, None, fB, 0, None, g.compute_likelihood(True), None, max_id, sg, True)
        assert np.allclose(so, sg, rtol=1e-8), (fA, fB, np.abs(so - sg).max())
    g.free_gpu()


def test_validation_step_scores_by_full_likelihoods(small_pyramid):
    """debug_step_max_likelihood (cuda_lib_gl.py:2109-2293): same trajectory as the oracle, and its full-likelihood scores
    agree with likelihood_t + delta of the production path (the reference's own check, :2196-2220)."""
    from graal_b200.sampler import CUR
    inp, o, g = _pair(small_pyramid, 2, 41)
    H.scramble(o, np.random.RandomState(6), 30, g)
    n = o.n_new_frags
    sched = np.random.RandomState(9).randint(0, n, size=6)
    for it, fA in enumerate(sched):
        ro = o.debug_step_max_likelihood(int(fA), 2)
        rg = g.debug_step_max_likelihood(int(fA), 2)
        assert o.score.dtype == g.score.dtype == np.float32
        assert np.allclose(o.score, g.score, rtol=3e-7), (it, np.abs(o.score - g.score).max())
        assert tuple(ro[1:]) == tuple(rg[1:]) and abs(float(ro[0]) - float(rg[0])) <= 3e-7 * abs(float(ro[0])), (it, ro, rg)
        assert H.slots_diff(o.cur, g.slot_to_host(CUR)) == [], it
    # full(candidate) vs likelihood_t + delta(candidate): the device gives the oracle's value for both, and the identity holds
    # wherever it holds in the reference (it does not for every translocation: the delta kernel only revisits the pixels of
    # the two touched contigs of the CURRENT structure, kernels3.cu:3225-3249)
    from graal_b200.sampler import CAND0
    from oracle import likelihood as L
    fA = int(sched[0])
    max_id = int(g.modify_gl_cuda_buffer(fA)); assert max_id == int(o.modify_gl_cuda_buffer(fA))
    lt, lo = g.eval_likelihood(), o.eval_likelihood()
    nb = g.return_neighbours(fA, 2); assert nb == o.return_neighbours(fA, 2)
    g.score = np.zeros(13 * len(nb)); o.score = np.zeros(13 * len(nb)); o.delta = np.zeros(13 * len(nb))
    held = 0
    for x, fB in enumerate(nb):
        g.stream_likelihood(fA, None, None, fB, x, lt, max_id)
        o.stream_likelihood(fA, fB, x, lo, max_id)
        full_g = np.array([g.full_likelihood_of_slot(CAND0 + j) for j in range(13)])
        full_o = np.array([L.evaluate_likelihood(o.ws.collector[j], o.lv, o.param_simu).sum() for j in range(13)])
        sl = slice(13 * x, 13 * x + 13)
        assert np.allclose(full_g, full_o, rtol=1e-8), (fB, np.abs(full_g - full_o).max())
        assert np.allclose(g.score[sl], o.score[sl], rtol=1e-8), (fB, np.abs(g.score[sl] - o.score[sl]).max())
        ok = np.abs(full_o - o.score[sl]) < 1e-2
        assert np.allclose(full_g[ok], g.score[sl][ok], rtol=1e-8), fB
        held += int(ok.sum())
    assert held >= 9 * len(nb)
    g.free_gpu()


def test_older_proposal_rule_and_its_step(small_pyramid):
    """define_neighbourhood / old_return_neighbours / step_max_likelihood_4_visu (cuda_lib_gl.py:2548-2561, 2333-2360,
    3140-3323) against the oracle, same RandomState."""
    from graal_b200.sampler import CUR
    inp, o, g = _pair(small_pyramid, 2, 51)
    o.define_neighbourhood(); g.define_neighbourhood()
    for i in range(o.n_frags):
        assert np.array_equal(o.sorted_neighbours[i], g.sorted_neighbours[i]), i
    H.scramble(o, np.random.RandomState(2), 30, g)
    n = o.n_new_frags
    sched = np.random.RandomState(3).randint(0, n, size=12)
    for it, fA in enumerate(sched):
        assert o.old_return_neighbours(int(fA), 3) == g.old_return_neighbours(int(fA), 3)
        ro = o.step_max_likelihood_4_visu(int(fA), 3)
        rg = g.step_max_likelihood_4_visu(int(fA), 3)
        assert tuple(ro[1:7]) == tuple(rg[1:7]) and ro[8] == rg[8], (it, ro, rg)
        assert abs(ro[0] - rg[0]) <= 1e-7 * abs(ro[0]) and abs(ro[7] - rg[7]) < 1e-12, (it, ro, rg)
        assert H.slots_diff(o.cur, g.slot_to_host(CUR)) == [], it
    g.free_gpu()


def test_local_flip_and_scramblers(small_pyramid):
    """local_flip (cuda_lib_gl.py:1056-1154), modify_genome (:1521-1537), diagnosis (:1016-1042)."""
    from graal_b200.sampler import CUR, CAND0
    inp, o, g = _pair(small_pyramid, 2, 61)
    o.rng = np.random.RandomState(5); g.rng = np.random.RandomState(5)
    o.modify_genome(15); g.modify_genome(15)
    assert H.slots_diff(o.cur, g.slot_to_host(CUR)) == []
    assert g.diagnosis(g.gpu_vect_frags, 0, 0, 0) == []
    max_id = int(o.modify_gl_cuda_buffer()); g.modify_gl_cuda_buffer()
    rng = np.random.RandomState(12)
    for case in range(8):
        fA, mode = int(rng.randint(o.n_new_frags)), int(rng.choice([12, 13, 14]))
        ref = o.local_flip(fA, mode, max_id)
        got = g.local_flip(fA, mode, max_id)
        assert H.slots_diff(ref, got) == [], (case, fA, mode)
        if mode < 13:
            assert H.slots_diff(ref, g.slot_to_host(CAND0 + mode)) == []
    g.free_gpu()


def test_insert_repeats_on_a_level_with_duplicated_bins(small_pyramid):
    """insert_repeats (cuda_lib_gl.py:1512-1519)."""
    from graal_b200.sampler import CUR
    inp, o, g = _pair(small_pyramid, 2, 71, allow_repeats=True)
    if not np.any(o.cur["rep"] == 1):
        pytest.skip("no duplicated bin on this level")
    o.insert_repeats(3); g.insert_repeats(3)
    assert H.slots_diff(o.cur, g.slot_to_host(CUR)) == []
    g.free_gpu()
