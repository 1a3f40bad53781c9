"""The sampler surface end to end: identical accepted-move trajectories under identical RNG draws."""
import os

import numpy as np
import pytest

from graal_b200.driver import start_EM, replay_simu, load_mutations
from graal_b200.level import prepare_sampler_inputs
from oracle import mutations as M
import helpers as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gpu_sampler(pyr, level, seed, **kw):
    from graal_b200.sampler import sampler
    inp = prepare_sampler_inputs(pyr, level, **kw)
    g = sampler.from_inputs(inp, rng=np.random.RandomState(seed))
    p, dm = H.default_params(pyr)
    g.set_parameters(p, dm)
    return inp, g


def test_live_trajectory_vs_oracle(small_pyramid):
    """Oracle and device driven by the same schedule and the same RandomState seed: every returned
    9-tuple agrees (integers exactly, scores to 1e-9 relative) and the genomes stay bit-identical."""
    from graal_b200.sampler import CUR
    inp, g = gpu_sampler(small_pyramid, 2, 77, blacklist_contigs=(5,))
    o = H.make_oracle(inp, small_pyramid, seed=77)
    rows_o, rows_g = [], []
    tr_o = start_EM(o, 3, 3, scrambled=True, max_steps=150, on_step=lambda it, tr: rows_o.append(M.copy_slot(o.cur)) if it % 50 == 0 else None)
    tr_g = start_EM(g, 3, 3, scrambled=True, max_steps=150, on_step=lambda it, tr: rows_g.append(g.slot_to_host(CUR)) if it % 50 == 0 else None)
    assert np.array_equal(tr_o.mutations(), tr_g.mutations())
    assert tr_o.n_contigs == tr_g.n_contigs
    assert np.allclose(tr_o.likelihood, tr_g.likelihood, rtol=1e-7, atol=0)
    assert np.allclose(tr_o.mean_len, tr_g.mean_len, rtol=1e-12) and np.allclose(tr_o.dist_from_init_genome, tr_g.dist_from_init_genome, rtol=0, atol=1e-12)
    for a, b in zip(rows_o, rows_g):
        assert H.slots_diff(a, b) == []
    assert -1 in tr_g.op_sampled                               # blacklisted bins are skipped (op = -1)
    g.free_gpu()


def test_golden_trajectory_10k_steps(yeast_pyramid, tmp_path):
    """BASELINE: 'under identical RNG draws the accepted-move trajectory must match for the first 10^4
    steps' -- against the frozen oracle trajectory of tests/golden/traj_c1_l3.npz."""
    from graal_b200.sampler import CUR
    z = np.load(os.path.join(GOLD, "traj_c1_l3.npz"))
    inp, g = gpu_sampler(yeast_pyramid, int(z["level"]), int(z["seed"]))
    n_steps = z["mutations"].shape[0]
    tr = start_EM(g, n_steps // g.n_new_frags + 1, 3, scrambled=True, max_steps=n_steps)
    got = tr.mutations()
    same = np.all(got == z["mutations"], axis=1)
    first_bad = int(np.argmin(same)) if not same.all() else -1
    margin = z["margins"][first_bad] if first_bad >= 0 else None
    assert first_bad < 0, "diverged at step %d (draw-to-boundary margin %s)" % (first_bad, margin)
    assert np.allclose(tr.likelihood, z["likelihood"], rtol=1e-7, atol=0)
    assert np.array_equal(np.array(tr.n_contigs), z["n_contigs"])
    final = g.slot_to_host(CUR)
    for k in M.FIELDS:
        assert np.array_equal(final[k], z["state_" + k]), k
    # export the trace (main_gl.py:321-342) and rebuild the genome from it (replay_simu :140-207)
    tr.save_behaviour_to_txt(str(tmp_path))
    muts = load_mutations(os.path.join(str(tmp_path), "list_mutations.txt"))
    assert len(muts) == n_steps
    inp2, g2 = gpu_sampler(yeast_pyramid, int(z["level"]), 0)
    replay_simu(g2, muts, scrambled=True)
    again = g2.slot_to_host(CUR)
    g.modify_gl_cuda_buffer(); g2.modify_gl_cuda_buffer()
    assert H.slots_diff(g.slot_to_host(CUR), g2.slot_to_host(CUR)) == []
    g.free_gpu(); g2.free_gpu()


def test_golden_trajectory_level1_one_cycle(yeast_pyramid):
    """The same identity on the largest C1 level (1,672 bins, 5,000 sub-frags -- the size the original kernels are
    benchmarked on): one whole cycle from the exploded genome, against the frozen oracle trajectory."""
    from graal_b200.sampler import CUR
    path = os.path.join(GOLD, "traj_c1_l1.npz")
    if not os.path.exists(path):
        pytest.skip("fixture not generated (python tests/golden/make_golden.py traj1, about 40 minutes)")
    z = np.load(path)
    inp, g = gpu_sampler(yeast_pyramid, int(z["level"]), int(z["seed"]))
    n_steps = z["mutations"].shape[0]
    tr = start_EM(g, n_steps // g.n_new_frags + 1, 3, scrambled=True, max_steps=n_steps)
    got = tr.mutations()
    same = np.all(got == z["mutations"], axis=1)
    first_bad = int(np.argmin(same)) if not same.all() else -1
    assert first_bad < 0, "diverged at step %d (draw-to-boundary margin %s)" % (first_bad, z["margins"][first_bad])
    assert np.allclose(tr.likelihood, z["likelihood"], rtol=1e-7, atol=0)
    assert np.array_equal(np.array(tr.n_contigs), z["n_contigs"])
    final = g.slot_to_host(CUR)
    for k in M.FIELDS:
        assert np.array_equal(final[k], z["state_" + k]), k
    g.free_gpu()


def test_golden_trajectory_with_nuisance_parameters(yeast_pyramid):
    z = np.load(os.path.join(GOLD, "traj_c1_l2_nuis.npz"))
    inp, g = gpu_sampler(yeast_pyramid, int(z["level"]), int(z["seed"]))
    g.bins = np.arange(10.0, 510.0, 10.0)
    n_steps = z["mutations"].shape[0]
    tr = start_EM(g, n_steps // g.n_new_frags + 1, 3, sample_param=True, scrambled=True, max_steps=n_steps)
    assert np.array_equal(tr.mutations(), z["mutations"]), int(np.argmin(np.all(tr.mutations() == z["mutations"], axis=1)))
    assert np.array_equal(np.array(tr.success), z["success"])
    for k in ("fact", "slope", "d_max", "d_nuc"):
        assert np.allclose(np.array(getattr(tr, k), dtype=np.float64), z[k], rtol=1e-6), k
    assert np.allclose(tr.likelihood, z["likelihood"], rtol=1e-7, atol=0)
    g.free_gpu()


def test_surface_attributes(small_pyramid):
    inp, g = gpu_sampler(small_pyramid, 2, 1)
    assert g.n_tmp_struct == 13 and len(g.modification_str) >= 13 and int(g.n_new_frags) == inp.n_new_frags
    g.setup_texture()
    g.init_likelihood()
    assert np.isfinite(g.likelihood_t)
    g.gpu_vect_frags.copy_from_gpu()
    assert np.array_equal(g.gpu_vect_frags.pos, inp.S_o_A_frags["pos"]) and np.all(g.gpu_vect_frags.ori == 1)
    full_order, content = g.genome_content()
    assert len(full_order) == inp.n_new_frags and sum(len(v["id"]) for v in content.values()) == inp.n_new_frags
    fo, dc, fo_high = g.display_current_matrix(None)
    assert [int(x) for x in fo] == [int(x) for x in full_order] and len(fo_high) == inp.init_n_sub_frags
    assert g.gpu_launches > 0
    g.free_gpu()


def test_repeat_levels(yeast_pyramid):
    """H3: levels with duplicated ("repeat") bins -- the expected value of a pixel is the float32 sum
    over ACTIVE copy pairs; unique x unique pairs stay on the sparse kernels, every pixel touching a
    duplicated bin goes through the per-pixel repeat path.  Full likelihood, the 13 deltas (fA a
    unique bin, an original of a duplicated bin, a repeat copy) and a live trajectory vs the oracle."""
    from graal_b200.sampler import sampler, CUR, CAND0
    from oracle import likelihood as L
    inp = prepare_sampler_inputs(yeast_pyramid, 3, allow_repeats=True)
    assert inp.n_new_frags > inp.n_frags
    g = sampler.from_inputs(inp, rng=np.random.RandomState(5))
    p, dm = H.default_params(yeast_pyramid)
    g.set_parameters(p, dm)
    o = H.make_oracle(inp, yeast_pyramid, seed=5)
    rng = np.random.RandomState(12)
    H.scramble(o, rng, 40, g)
    assert H.slots_diff(o.cur, g.slot_to_host(CUR)) == []
    max_id = o.modify_gl_cuda_buffer(); g.modify_gl_cuda_buffer()
    fo, fg = o.eval_likelihood(), g.eval_likelihood()
    assert abs(fo - fg) <= 1e-7 * abs(fo)
    n, N = o.n_new_frags, o.n_frags
    dup = int(inp.id_frag_duplicated[0])
    copy = int(np.nonzero(np.asarray(inp.S_o_A_frags["rep"]) == 1)[0][0])
    for fA, fB in ((3, 60), (dup, 17), (copy, 100), (50, copy), (copy, dup)):
        M.perform_modifications(o.ws, o.cur, fA, fB, max_id)
        no_rep, rep = o.candidate_index_sets(fA, fB)
        bi, bj, dg, glob = L.delta_pixels(o.lv, no_rep, rep, o.uniq_frags)
        g.score_neighbours(fA, [fB])
        got = g._fetch()[16:29].copy()
        for j in range(13):
            assert H.slots_diff(o.ws.collector[j], g.slot_to_host(CAND0 + j)) == [], (fA, fB, j)
            new = L.pixel_loglik(o.ws.collector[j], o.lv, o.param_simu, bi, bj, dg)
            old = o.curr_likelihood[glob]
            ref, mass = float(np.sum(new - old)), float(np.abs(new).sum() + np.abs(old).sum())
            assert abs(got[j] - ref) <= 1e-6 * abs(ref) + H.MASS_FLOOR * mass + 1e-9, (fA, fB, j, got[j], ref)
    # de-activate a copy (mode 8) on both sides, then a live trajectory
    M.apply_mutation(o.ws, o.cur, copy, 0, 8, max_id, o.id_contigs)
    g.test_copy_struct(copy, 0, 8, int(max_id))
    assert H.slots_diff(o.cur, g.slot_to_host(CUR)) == [] and g.slot_to_host(CUR)["activ"][copy] == 0
    fo, fg = o.eval_likelihood(), g.eval_likelihood()
    assert abs(fo - fg) <= 1e-7 * abs(fo)
    o.rng = np.random.RandomState(3); g.rng = np.random.RandomState(3)
    for it in range(40):
        fA = int(np.random.RandomState(100 + it).randint(n))
        ro, rg = o.step_max_likelihood(fA, 3), g.step_max_likelihood(fA, 3)
        assert (ro[1], ro[5], ro[6]) == (rg[1], rg[5], rg[6]), (it, ro, rg)
        assert abs(ro[0] - rg[0]) <= 1e-7 * abs(ro[0]) and abs(ro[7] - rg[7]) < 1e-12
    assert H.slots_diff(o.cur, g.slot_to_host(CUR)) == []
    g.free_gpu()


def test_lanes_and_graphs_match_the_plain_abi(small_pyramid):
    """graal_score_proposal (lanes, captured graphs, replayed with new bins) against the unfused serial calls
    graal_build_candidates + graal_delta_loglik on the same state: the 13 deltas are the same doubles, and
    graal_dist_candidates equals graal_dist_genome of each candidate slot."""
    import ctypes as C
    from graal_b200.sampler import CUR, CAND0, N_LANES, OFF_DIST, N_TMP_STRUCT
    from graal_b200._lib import check
    inp, g = gpu_sampler(small_pyramid, 2, 5)
    rng = np.random.RandomState(9)
    n = int(g.n_new_frags)
    lib, ctx = g.lib, g.ctx
    for rnd in range(4):                                    # rounds 1.. replay the graphs captured in round 0
        g.modify_gl_cuda_buffer()
        fA = int(rng.randint(n))
        fBs = [int(x) for x in rng.choice(np.setdiff1d(np.arange(n), [fA]), 3, replace=False)]
        g.score_neighbours(fA, fBs, with_dist=True)
        fast = g._fetch().copy()
        for x, fB in enumerate(fBs):
            check(lib.graal_build_candidates(ctx, CUR, CAND0, fA, fB, -1, 0x1FFF))
            check(lib.graal_delta_loglik(ctx, CUR, CAND0, N_TMP_STRUCT, fA, fB, -1, g._ptr(g.d_out, 32)))
            ref = g._fetch()[32:32 + N_TMP_STRUCT].copy()
            got = fast[16 + N_TMP_STRUCT * x: 16 + N_TMP_STRUCT * (x + 1)]
            # the fused path copies candidate 8 from candidate 0 and scores candidates 3, 5, 7 as second-level deltas
            # against 2, 4, 6 (same terms, summed in another order); everything else is the same arithmetic
            same = np.array([0, 1, 2, 4, 6, 9, 10, 11, 12])
            assert np.array_equal(got[same], ref[same]), (rnd, x, got, ref)
            assert got[8] == got[0]
            paired = np.array([3, 5, 7])
            assert np.allclose(got[paired], ref[paired], rtol=1e-9, atol=1e-7), (rnd, x, got[paired] - ref[paired])
            for k in (0, 5, 12):
                check(lib.graal_dist_genome(ctx, CAND0 + k, g._ptr(g.d_init_prev), g._ptr(g.d_init_next),
                                            g._ptr(g.d_init_orientable), g._ptr(g.d_dist_skip), g._ptr(g.d_out, 60)))
                assert g._fetch()[60] == fast[OFF_DIST + N_TMP_STRUCT * x + k]
        # move on: commit something so that the next round sees another state
        check(lib.graal_commit_scored(ctx, CUR, CAND0, fA, fBs[1], -1, int(rng.randint(13)), 1))
    g.free_gpu()


def test_incremental_likelihood_mode(small_pyramid):
    """incremental_likelihood=True: the likelihood of the current state is carried from the committed
    candidate's score instead of being recomputed every step (resync every few steps here).  Same accepted
    moves, likelihood trace within the full-likelihood tolerance."""
    inp, g = gpu_sampler(small_pyramid, 2, 77)
    inp2, h = gpu_sampler(small_pyramid, 2, 77)
    h.incremental_likelihood = True
    h.incremental_resync = 7
    tr_g = start_EM(g, 3, 3, scrambled=True, max_steps=120)
    tr_h = start_EM(h, 3, 3, scrambled=True, max_steps=120)
    assert np.array_equal(tr_g.mutations(), tr_h.mutations())
    # full(t) + delta against full(t + 1).  The reference's delta is not the exact difference of two full
    # likelihoods (diagonal pixels are never re-scored, quirk Q4; its own cross-check cuda_lib_gl.py:2196-2220 is
    # approximate), so the carried value drifts between resyncs: measured worst 1e-4 relative on this run.
    a, b = np.array(tr_g.likelihood), np.array(tr_h.likelihood)
    worst = float(np.max(np.abs(a - b) / np.abs(a)))
    assert worst <= 1e-3, worst
    assert np.array_equal(np.array(tr_g.dist_from_init_genome), np.array(tr_h.dist_from_init_genome))
    assert h.gpu_launches < g.gpu_launches
    g.free_gpu(); h.free_gpu()


def test_device_contact_lists_match_the_host_preparation(small_pyramid):
    """graal_coo_to_lists (device: keys, radix sort, duplicate sums, row prefix sum) vs build_contact_lists (host NumPy,
    the restatement of cuda_lib_gl.py:153-172), also with entries given in both triangles, duplicated, on the diagonal and
    with zero counts."""
    from graal_b200.sampler import build_contact_lists
    inp, g = gpu_sampler(small_pyramid, 1, 1)
    W = inp.init_n_sub_frags
    r, c, v = (np.asarray(a) for a in inp.sub_coo)
    rp_h, ct_h = build_contact_lists(inp.sub_coo, W)
    assert np.array_equal(g.d_rowptr.cpu().numpy(), rp_h) and np.array_equal(g.d_contacts.cpu().numpy(), ct_h)
    rng = np.random.RandomState(2)
    k = rng.choice(r.shape[0], 500, replace=False)
    r2 = np.concatenate([r, c[k], np.arange(20), r[:50]])            # lower-triangle copies, diagonal entries, zero counts
    c2 = np.concatenate([c, r[k], np.arange(20), c[:50]])
    v2 = np.concatenate([v, v[k], np.full(20, 7), np.zeros(50)]).astype(np.float32)
    rp_d, ct_d = g.device_contact_lists((r2, c2, v2), W)
    dense = np.zeros((W, W), dtype=np.float64)
    np.add.at(dense, (np.minimum(r2, c2), np.maximum(r2, c2)), v2)
    np.fill_diagonal(dense, 0.0)
    rr, cc = np.nonzero(dense)
    ct = ct_d.cpu().numpy()
    assert np.array_equal(rp_d.cpu().numpy(), np.r_[0, np.cumsum(np.bincount(rr, minlength=W))])
    assert np.array_equal(ct[:, 0], cc) and np.array_equal(ct[:, 1].copy().view(np.float32), dense[rr, cc].astype(np.float32))
    g.free_gpu()


def test_fused_prologue_matches_the_general_sequences(small_pyramid):
    """graal_stats_relabel: the three-launch prologue (first step: general path + seeded bound; then fused) against the
    oracle's statistics and relabel over a run of committed moves, and against the general path (GRAAL_FUSED_PROLOGUE=0)."""
    import os
    from graal_b200.sampler import CUR
    inp, g = gpu_sampler(small_pyramid, 2, 3)
    os.environ["GRAAL_FUSED_PROLOGUE"] = "0"
    try:
        inp2, g2 = gpu_sampler(small_pyramid, 2, 3)
    finally:
        os.environ.pop("GRAAL_FUSED_PROLOGUE")
    o = H.make_oracle(inp, small_pyramid, seed=3)
    rng = np.random.RandomState(1)
    n = o.n_new_frags
    for it in range(40):
        fA = int(rng.randint(n))
        ro, r1, r2 = o.step_max_likelihood(fA, 3), g.step_max_likelihood(fA, 3), g2.step_max_likelihood(fA, 3)
        assert tuple(ro[1:7]) == tuple(r1[1:7]) == tuple(r2[1:7]), (it, ro, r1, r2)
        assert r1[0] == r2[0]
        assert H.slots_diff(o.cur, g.slot_to_host(CUR)) == [] and H.slots_diff(o.cur, g2.slot_to_host(CUR)) == []
    # the device-resident replay (no fetch at all) seeds its bound by itself
    g.slot_from_host(CUR, g2.slot_to_host(CUR))
    for it in range(5):
        fA = int(rng.randint(n))
        nb = g2.return_neighbours(fA, 3); nb.sort()
        if not nb:
            continue
        g.step_device(fA, nb, nb[0], 6); g2.step_device(fA, nb, nb[0], 6)
        assert H.slots_diff(g.slot_to_host(CUR), g2.slot_to_host(CUR)) == []
    g.free_gpu(); g2.free_gpu()


def test_device_draw_is_the_host_draw(small_pyramid):
    """The candidate draw made on the device (k_draw_candidates, right behind the scores) against the host draw
    (sampler._sample, pinned to the reference's lines by tests/test_reference_host_logic.py): same candidate, same number of
    candidates left, same normalised weights bit for bit, on thousands of score vectors -- and the same steps, tuples, genome and
    RandomState position when whole steps run either way."""
    import ctypes as C
    import types
    import torch
    from graal_b200 import _lib
    from graal_b200.sampler import sampler, CUR, OFF_SUB, OFF_DRAW
    inp, g = gpu_sampler(small_pyramid, 2, 3)
    gen = np.random.RandomState(17)
    checked_draws = 0
    for case in range(1500):
        n_nb = int(gen.randint(1, 17))
        n = 13 * n_nb
        delta = gen.randn(n) * gen.choice([0.01, 1.0, 5.0, 40.0, 400.0])
        lt = -1e6 * gen.rand()
        if case % 7 == 0:
            delta[gen.randint(n)] += 100.0
        if case % 11 == 0:
            delta[:] = delta[0]
        if case % 97 == 0:
            delta[gen.randint(n)] = np.nan
        u = float(gen.rand())
        nb = gen.randint(0, int(g.n_new_frags), size=n_nb).astype(np.int32)
        g.d_out[16:16 + n] = torch.from_numpy(delta).to(g.device)
        g.d_out[0] = lt
        fbs = (C.c_int32 * n_nb)(*[int(x) for x in nb])
        _lib.check(g.lib.graal_draw_candidates(g.ctx, g._p_out + 8 * 16, g._p_out, 0.0, 0, n_nb, fbs, u, g.d_sel.data_ptr(),
                                                g._p_out + 8 * OFF_SUB, g._p_out + 8 * OFF_DRAW))
        out = g._fetch()
        sel = g.d_sel.cpu().numpy()
        score = delta + np.float64(lt)
        stub = types.SimpleNamespace(rng=types.SimpleNamespace(random_sample=lambda: u), _remove_cache={}, sub_score=None, _fast_weights=True)
        if np.isnan(score).any():
            assert sel[2] == 1 and out[OFF_DRAW + 3] == 1.0, case
            continue
        want = sampler._sample(stub, score.copy(), 1.0)
        n_ok = len(stub.sub_score)
        assert sel[2] == 0 and sel[0] == want == int(out[OFF_DRAW + 1]), (case, sel, want)
        assert sel[1] == n_ok == int(out[OFF_DRAW + 2]), (case, sel, n_ok)
        assert sel[3] == nb[want // 13] and sel[4] == want % 13
        assert np.array_equal(out[OFF_SUB:OFF_SUB + n_ok], stub.sub_score), case
        assert out[OFF_DRAW] == score[want]
        checked_draws += int(n_ok > 1)
    assert checked_draws > 500
    g.free_gpu()
    # whole steps either way
    runs = []
    for dev in (False, True):
        inp, g = gpu_sampler(small_pyramid, 2, 29)
        g.device_draw = dev
        H.scramble(H.make_oracle(inp, small_pyramid, seed=1), np.random.RandomState(4), 25, g)
        sched = np.random.RandomState(6).randint(0, int(g.n_new_frags), size=60)
        tuples = [g.step_max_likelihood(int(fA), 3) for fA in sched]
        runs.append((tuples, g.slot_to_host(CUR), g.rng.rand(), [np.array(g.sub_score)]))
        g.free_gpu()
    (ta, sa, ra, wa), (tb, sb, rb, wb) = runs
    assert ra == rb                                                  # the stream advanced by the same number of draws
    assert H.slots_diff(sa, sb) == []
    for a, b in zip(ta, tb):
        assert tuple(a[1:6]) == tuple(b[1:6]) and a[6] == b[6] and a[0] == b[0] and a[7] == b[7], (a, b)
    assert np.array_equal(wa[0], wb[0])
