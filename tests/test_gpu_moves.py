"""Device structure kernels vs the oracle: BIT-EXACT (integer state)."""
import os

import numpy as np
import pytest

from graal_b200.level import prepare_sampler_inputs
from oracle import mutations as M
import helpers as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def pair(small_pyramid):
    from graal_b200.sampler import sampler
    inp = prepare_sampler_inputs(small_pyramid, 1)
    o = H.make_oracle(inp, small_pyramid)
    g = sampler.from_inputs(inp, rng=np.random.RandomState(1000))
    p, dm = H.default_params(small_pyramid)
    g.set_parameters(p, dm)
    yield o, g
    g.free_gpu()


def test_single_kernels_bit_exact(pair):
    """Each mutation kernel on its own, from random scrambled states (incl. circular contigs),
    into a destination slot with known content (persistent-slot semantics)."""
    from graal_b200.sampler import CUR, CAND0
    o, g = pair
    rng = np.random.RandomState(21)
    n = o.n_new_frags
    for rnd in range(12):
        H.scramble(o, rng, 15, g)
        max_id = int(o.modify_gl_cuda_buffer()); g.modify_gl_cuda_buffer()
        assert H.slots_diff(o.cur, g.slot_to_host(CUR)) == []
        fA, fB = (int(x) for x in rng.choice(n, 2, replace=False))
        sentinel = M.copy_slot(o.cur)
        sentinel["pos"][:] = 12345
        cases = [("FLIP", lambda d: M.flip_frag(d, o.cur, fA), 0),
                 ("SWAP_ACTIV", lambda d: M.swap_activity_frag(d, o.cur, fA, max_id), 0),
                 ("COPY", lambda d: M.simple_copy(d, o.cur), 0)]
        for up in (0, 1):
            cases.append(("SPLIT", lambda d, up=up: M.split_contig(d, o.cur, np.zeros(n, np.int32), fA, up, max_id), up))
        cases.append(("PASTE", lambda d: M.paste_contigs(d, o.cur, fA, fB, max_id), 0))
        for name, fn, aux in cases:
            exp = M.copy_slot(sentinel)
            fn(exp)
            g.slot_from_host(CAND0, sentinel)
            mx = g.apply_move(CUR, CAND0, name, fA, fB, aux, max_id)
            got = g.slot_to_host(CAND0)
            assert H.slots_diff(exp, got) == [], (name, aux, fA, fB)
            assert mx == int(got["id_c"].max())
        # pop_out then the four insertions, both orientations
        pop = M.copy_slot(sentinel); pid = np.zeros(n, np.int32)
        M.pop_out_frag(pop, o.cur, pid, fA, max_id)
        g.slot_from_host(CAND0, sentinel)
        mx2 = g.apply_move(CUR, CAND0, "POP_OUT", fA, 0, 0, max_id)
        assert H.slots_diff(pop, g.slot_to_host(CAND0)) == [] and mx2 == int(pid.max())
        for k, fn in ((1, M.pop_in_frag_1), (2, M.pop_in_frag_2), (3, M.pop_in_frag_3), (4, M.pop_in_frag_4)):
            for ori in (1, -1):
                exp = M.copy_slot(sentinel)
                fn(exp, pop, fA, fB, mx2, ori)
                g.slot_from_host(CAND0 + 1, sentinel)
                g.apply_move(CAND0, CAND0 + 1, "POP_IN_%d" % k, fA, fB, ori, mx2)
                assert H.slots_diff(exp, g.slot_to_host(CAND0 + 1)) == [], (k, ori, fA, fB)


def test_fused_candidates_bit_exact(pair):
    from graal_b200.sampler import CUR, CAND0
    o, g = pair
    rng = np.random.RandomState(22)
    n = o.n_new_frags
    for rnd in range(25):
        H.scramble(o, rng, 5, g)
        max_id = o.modify_gl_cuda_buffer(); g.modify_gl_cuda_buffer()
        fA, fB = (int(x) for x in rng.choice(n, 2, replace=False))
        M.perform_modifications(o.ws, o.cur, fA, fB, max_id)
        g.perform_modifications(fA, fB)
        for j in range(13):
            assert H.slots_diff(o.ws.collector[j], g.slot_to_host(CAND0 + j)) == [], (rnd, j, fA, fB)
        # explicit max_id and a partial mask leave the other collector slots untouched
        before = [g.slot_to_host(CAND0 + j) for j in range(13)]
        fA2, fB2 = (int(x) for x in rng.choice(n, 2, replace=False))
        g.perform_modifications(fA2, fB2, int(max_id), mask=(1 << 4) | (1 << 11))
        M.pop_out_pop_in(o.ws, o.cur, fA2, fB2, 4, max_id)
        M.transloc(o.ws, o.cur, fA2, fB2, max_id)
        for j in range(13):
            got = g.slot_to_host(CAND0 + j)
            if j in (4, 11):
                assert H.slots_diff(o.ws.collector[j], got) == []
            else:
                assert H.slots_diff(before[j], got) == []
        for j in (9, 10, 12):       # bring the oracle's persistent slots back in line with the device's
            o.ws.collector[j] = {k: before[j][k].copy() for k in M.FIELDS}


def test_degenerate_proposal_keeps_persistent_slots(pair):
    """id_fB == id_fA: paste_contigs writes nothing for the contig's bins."""
    from graal_b200.sampler import CAND0
    o, g = pair
    max_id = o.modify_gl_cuda_buffer(); g.modify_gl_cuda_buffer()
    for j in range(13):
        g.slot_from_host(CAND0 + j, o.ws.collector[j])
    big = int(np.argmax(o.cur["l_cont"]))
    M.perform_modifications(o.ws, o.cur, big, big, max_id)
    g.perform_modifications(big, big)
    for j in range(13):
        assert H.slots_diff(o.ws.collector[j], g.slot_to_host(CAND0 + j)) == [], j


def test_relabel_commit_and_explode(pair):
    from graal_b200.sampler import CUR
    o, g = pair
    rng = np.random.RandomState(23)
    H.scramble(o, rng, 30, g)
    assert int(o.modify_gl_cuda_buffer()) == int(g.modify_gl_cuda_buffer())
    assert H.slots_diff(o.cur, g.slot_to_host(CUR)) == []
    o.explode_genome(); g.explode_genome()
    got = g.slot_to_host(CUR)
    assert H.slots_diff(o.cur, got) == []
    assert np.all(got["l_cont"] == 1)
    assert M.check_invariants(got) == []
    # stats of an exploded genome
    g.lib.graal_state_stats(g.ctx, CUR, g._ptr(g.d_out, 4))
    out = g._fetch()
    assert int(out[4]) == o.n_new_frags and int(out[5]) == 1 and int(out[7]) == 1
    assert abs(out[6] - o.cur["l_cont_bp"].mean()) < 1e-9


@pytest.mark.parametrize("level", [2, 3])
def test_golden_candidates(yeast_pyramid, level):
    from graal_b200.sampler import sampler, CUR, CAND0
    z = np.load(os.path.join(GOLD, "like_c1_l%d.npz" % level))
    inp = prepare_sampler_inputs(yeast_pyramid, level)
    g = sampler.from_inputs(inp)
    g.slot_from_host(CUR, {k: z["state_" + k] for k in M.FIELDS})
    for (fA, fB), cands in zip(z["pairs"], z["candidates"]):
        g.perform_modifications(int(fA), int(fB), int(z["max_id"]))
        for j in range(13):
            got = g.slot_to_host(CAND0 + j)
            for fi, k in enumerate(M.FIELDS):
                assert np.array_equal(got[k], cands[j, fi]), (fA, fB, j, k)
    g.free_gpu()


def test_abi_errors(pair):
    from graal_b200.sampler import GraalError
    o, g = pair
    with pytest.raises(GraalError):
        g.apply_move(0, 99, "FLIP", 0)
    with pytest.raises(GraalError):
        g.apply_move(0, 0, "FLIP", 0)
    with pytest.raises(GraalError):
        g.perform_modifications(10 ** 9, 0)
    assert b"out of range" in g.lib.graal_last_error()
