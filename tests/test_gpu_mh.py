"""The Metropolis-Hastings / multiple-try variants of the sampler (graal_b200/mh.py) on the device against the
oracle's restatement (oracle/sampler.py), same RandomState: same returned tuples, same genomes."""
import numpy as np
import pytest

from graal_b200.level import prepare_sampler_inputs
import helpers as H

pytestmark = pytest.mark.gpu


def _pair(pyr, level, seed):
    from graal_b200.sampler import sampler
    inp = prepare_sampler_inputs(pyr, level)
    o = H.make_oracle(inp, pyr, seed=seed)
    g = sampler.from_inputs(inp, rng=np.random.RandomState(seed))
    p, dm = H.default_params(pyr)
    g.set_parameters(p, dm)
    return inp, o, g


def test_jump_sets_and_mh_candidates(small_pyramid):
    from graal_b200.sampler import CUR, CAND0
    inp, o, g = _pair(small_pyramid, 2, 11)
    o.set_jumping_distributions_parameters(4); g.set_jumping_distributions_parameters(4)
    for i in range(o.n_frags):
        assert np.array_equal(o.jump_dictionnary[i]["frags"], g.jump_dictionnary[i]["frags"]), i
        assert np.array_equal(o.jump_dictionnary[i]["proba"], g.jump_dictionnary[i]["proba"], equal_nan=True), i
    rng = np.random.RandomState(3)
    H.scramble(o, rng, 30, g)
    max_id = int(o.modify_gl_cuda_buffer()); g.modify_gl_cuda_buffer()
    n = o.n_new_frags
    # candidate structures (incl. contig ends, where paste / translocation are real moves) and the scores of both directions
    ends = np.nonzero((o.cur["prev"] == -1) | (o.cur["next"] == -1))[0]
    pairs = [(int(ends[0]), int(ends[3])), (int(ends[1]), int(rng.randint(n))), tuple(int(x) for x in rng.choice(n, 2, replace=False))]
    for fA, fB in pairs:
        o.all_modifications_metropolis(fA, fB, max_id, True)
        g.all_modifications_metropolis(fA, fB, max_id, True)
        for j in range(13):
            assert H.slots_diff(o.ws.collector[j], g.slot_to_host(CAND0 + j)) == [], (fA, fB, j)
        so = o.compute_all_score_MH(fA, {fB, int(o.cur["next"][fB]) if o.cur["next"][fB] >= 0 else fB}, True)
        sg = g.compute_all_score_MH(fA, {fB, int(o.cur["next"][fB]) if o.cur["next"][fB] >= 0 else fB}, True)
        assert np.allclose(so, sg, rtol=1e-9, atol=1e-6), (fA, fB, np.abs(so - sg).max())
        assert o.detect_impossibility(fA, [fB], True) == g.detect_impossibility(fA, [fB], True)
    g.free_gpu()


@pytest.mark.parametrize("variant", ["step_metropolis_hastings_s_a", "step_mtm"])
def test_mh_trajectories(small_pyramid, variant):
    from graal_b200.sampler import CUR
    inp, o, g = _pair(small_pyramid, 2, 21)
    H.scramble(o, np.random.RandomState(2), 40, g)                  # a rearranged genome: proposals that repair it are accepted
    o.set_jumping_distributions_parameters(3); g.set_jumping_distributions_parameters(3)
    o.init_likelihood(); g.init_likelihood()
    n = o.n_new_frags
    sched = np.random.RandomState(5).randint(0, n, size=14)
    accepted = 0
    for it, fA in enumerate(sched):
        before = o.cur["id_c"].copy(), o.cur["pos"].copy()
        ro = getattr(o, variant)(int(fA)); rg = getattr(g, variant)(int(fA))
        assert H.slots_diff(o.cur, g.slot_to_host(CUR)) == [], (variant, it)
        assert abs(ro[0] - rg[0]) <= 1e-7 * abs(ro[0]) and tuple(ro[1:5]) == tuple(rg[1:5]) and abs(ro[6] - rg[6]) < 1e-12, (variant, it, ro, rg)
        accepted += int(not (np.array_equal(before[0], o.cur["id_c"]) and np.array_equal(before[1], o.cur["pos"])))
    assert accepted > 0
    g.free_gpu()
