"""Host-side logic of the product (CPU): level derivation, contact lists, proposal tables,
distance metric and the Rippe fit, each against the oracle's statement of the same reference code."""
import numpy as np
import pytest

from graal_b200 import rippe as opti
from graal_b200.level import prepare_sampler_inputs, build_synthetic_pyramid
from graal_b200.sampler import build_contact_lists, neighbour_tables, dist_inter_genome, rippe_c1
from oracle import mutations as M, likelihood as L
from oracle import sampler as OS
import helpers as H

F32, I32 = np.float32, np.int32


def test_binning_rule(small_pyramid):
    """subsample_data_set (pyramid_sparse.py:406-431): bins of `factor` consecutive fragments, the
    last bin of a contig shorter, contigs with < factor fragments kept 1:1; contacts conserved."""
    l0, l1 = small_pyramid.levels[0], small_pyramid.levels[1]
    for c in np.unique(l0.contig_id):
        n0 = int((l0.contig_id == c).sum())
        n1 = int((l1.contig_id == c).sum())
        assert n1 == (n0 if n0 < 3 else -(-n0 // 3))
    assert l1.n_accu.sum() == l0.n_frags and l1.n_accu.max() == 3
    assert int(l1.vals.sum()) == int(l0.vals.sum())
    assert np.all(l1.rows <= l1.cols)
    # bins are contiguous runs of the level below
    assert np.array_equal(l1.sub_low[1:], l1.sub_high[:-1] + 1) and l1.sub_low[0] == 0
    assert np.array_equal(l1.end_pos, l0.end_pos[l1.sub_high]) and np.array_equal(l1.start_pos, l0.start_pos[l1.sub_low])
    s = l1.S_o_A_frags
    heads = s["pos"] == 0
    assert np.all(s["start_bp"][heads] == 0) and np.all(s["prev"][heads] == -1)
    assert M.check_invariants({**{k: s[k] for k in s if k in M.FIELDS}, "ori": np.ones_like(s["pos"]),
                               "rep": np.zeros_like(s["pos"]), "activ": np.ones_like(s["pos"]), "id_d": s["id"]}) == []


def test_sampler_inputs(small_pyramid):
    inp = prepare_sampler_inputs(small_pyramid, 2)
    assert inp.init_n_sub_frags == small_pyramid.levels[1].n_frags
    cnt = inp.np_sub_frags_id[:, 3]
    assert cnt.min() >= 1 and cnt.max() <= 3 and cnt.sum() == inp.init_n_sub_frags
    # sub ids consecutive in bin order, lengths in kb as float32(len_bp)/1000
    flat = np.concatenate([inp.np_sub_frags_id[b, :cnt[b]] for b in range(inp.n_frags)])
    assert np.array_equal(flat, np.arange(inp.init_n_sub_frags))
    sub_len_bp = small_pyramid.levels[1].S_o_A_frags["len_bp"]
    assert inp.np_sub_frags_len_bp[5, 0] == F32(sub_len_bp[inp.np_sub_frags_id[5, 0]]) / F32(1000.0)
    assert inp.mean_squared_frags_per_bin == F32(small_pyramid.levels[1].n_accu.astype(F32).mean() ** 2)
    with pytest.raises(ValueError):
        prepare_sampler_inputs(small_pyramid, 0)


def test_contact_lists_match_dense(small_pyramid):
    inp = prepare_sampler_inputs(small_pyramid, 1, blacklist_contigs=(6,))
    assert len(inp.id_frags_blacklisted) > 0
    o = OS.OracleSampler(inp, np.random.RandomState(0))
    black_subs = []
    for f in inp.id_frags_blacklisted:
        da = inp.np_sub_frags_id[inp.S_o_A_frags["id_d"][f]]
        black_subs.extend(int(da[k]) for k in range(da[3]))
    rowptr, contacts = build_contact_lists(inp.sub_coo, inp.init_n_sub_frags, black_subs, inp.mean_value_trans)
    W = inp.init_n_sub_frags
    dense = np.zeros((W, W), dtype=F32)
    rows = np.repeat(np.arange(W), np.diff(rowptr))
    dense[rows, contacts[:, 0]] = contacts[:, 1].view(F32)
    assert np.all(rows < contacts[:, 0])
    assert np.array_equal(dense, np.triu(o.hic_matrix, 1))
    # rows sorted by column, no explicit zeros
    assert np.all(contacts[:, 1].view(F32) != 0)
    for r in (0, W // 2):
        seg = contacts[rowptr[r]:rowptr[r + 1], 0]
        assert np.all(np.diff(seg) > 0)


def test_neighbour_tables_match_dense_rule(small_pyramid):
    for level, bl in ((2, ()), (1, (6,))):
        inp = prepare_sampler_inputs(small_pyramid, level, blacklist_contigs=bl)
        o = OS.OracleSampler(inp, np.random.RandomState(0))
        black_bins = [int(inp.S_o_A_frags["id_d"][f]) for f in inp.id_frags_blacklisted]
        xk, pk = neighbour_tables(inp.level_coo, inp.n_frags, black_bins)
        for i in range(inp.n_frags):
            d = o.distri_frags[i]
            nz = d["pk"] != 0
            if nz.any():
                assert np.array_equal(xk[i][nz], d["xk"][nz]) and np.array_equal(pk[i], d["pk"]), i
            else:
                assert np.array_equal(xk[i], d["xk"]) and np.array_equal(pk[i], d["pk"]), i


def test_dist_inter_genome_matches_reference_loop(small_pyramid):
    inp = prepare_sampler_inputs(small_pyramid, 2)
    o = H.make_oracle(inp, small_pyramid)
    assert o.dist_inter_genome(o.cur) == 0.0
    rng = np.random.RandomState(4)
    for it in range(10):
        H.scramble(o, rng, 6)
        c = o.cur
        ref = o.dist_inter_genome(c)
        got = dist_inter_genome(c["prev"], c["next"], c["ori"], c["id_d"], o.np_init_prev, o.np_init_next,
                                o.np_init_ori, o.np_init_orientable, o.id_frags_blacklisted, o.is_repeat,
                                o.n_new_frags, o.n_frags_4_dist)
        assert abs(ref - got) < 1e-12 and 0.0 <= got <= 1.0


def test_rippe_fit_matches_oracle(small_pyramid):
    inp = prepare_sampler_inputs(small_pyramid, 1)
    o = OS.OracleSampler(inp, np.random.RandomState(0))
    s = inp.S_o_A_frags
    max_kb = s["l_cont_bp"][s["start_bp"] == 0].mean() / 1000.
    bin_kb = s["len_bp"].mean() / 1000.0
    o.estimate_parameters(max_kb, bin_kb)
    p, y = opti.estimate_param_rippe(o.mean_contacts, o.bins)
    p_o, y_o = OS.estimate_param_rippe(o.mean_contacts, o.bins)
    assert np.allclose(p, p_o, rtol=1e-12) and np.allclose(y, y_o, rtol=1e-12)
    assert opti.estimate_max_dist_intra(p, inp.mean_value_trans) == OS.estimate_max_dist_intra(p_o, inp.mean_value_trans)
    assert rippe_c1(p[0], p[1], p[2]) == o.param_simu["c1"]
    # nuisance-step quirk Q10: peval reads param[3] as the amplitude
    assert opti.peval(100.0, [1.0, 9.6, -1.5, 3.0, 50.0]) == OS.peval(100.0, [1.0, 9.6, -1.5, 3.0, 50.0])


def test_edge_levels():
    """Ragged input: contigs of 1 and 2 fragments, a bin with a single sub-frag."""
    pyr = build_synthetic_pyramid([50_000, 400, 900, 30_000], 40, 2, seed=3, cis_rowsum=50.0, v_inter=0.01)
    inp = prepare_sampler_inputs(pyr, 1)
    assert inp.np_sub_frags_id[:, 3].min() == 1
    o = H.make_oracle(inp, pyr)
    like = o.eval_likelihood()
    assert np.isfinite(like)
    rng = np.random.RandomState(0)
    H.scramble(o, rng, 40)
    assert M.check_invariants(o.cur) == []
    o.explode_genome()
    # every bin is its own contig; a bin that was already a singleton keeps its orientation
    # (pop_out_frag is the identity for l_cont == 1, kernels3.cu:545-561)
    assert np.all(o.cur["l_cont"] == 1) and np.all(o.cur["pos"] == 0) and np.all(o.cur["prev"] == -1)
    assert len(np.unique(o.cur["id_c"])) == o.n_new_frags


def test_candidate_weights_in_c_match_the_numpy_statements():
    """graal_candidate_weights (the float64 arithmetic of the candidate draw in C, NumPy's pairwise sums included) against the
    NumPy statements of sampler._sample: same candidate, same weights bit for bit, same stream position."""
    import types
    from graal_b200 import sampler as GS, _lib
    _lib.build()
    gen = np.random.RandomState(0)
    for case in range(3000):
        n_nb = int(gen.randint(1, 17))
        score = -1e6 + gen.randn(13 * n_nb) * gen.choice([1e-3, 0.01, 1.0, 5.0, 40.0, 300.0])
        if case % 7 == 0:
            score[gen.randint(score.size)] += 100.0
        if case % 11 == 0:
            score[:] = score[0]
        if case % 13 == 0:
            score[gen.randint(score.size)] = np.nan if case % 2 else np.inf
        seed = int(gen.randint(1 << 30))
        a = types.SimpleNamespace(rng=np.random.RandomState(seed), _remove_cache={}, sub_score=None, _fast_weights=False)
        b = types.SimpleNamespace(rng=np.random.RandomState(seed), _remove_cache={}, sub_score=None, _fast_weights=True)
        try:
            ra = GS.sampler._sample(a, score.copy(), 1.0)
        except ValueError:
            with pytest.raises(ValueError):
                GS.sampler._sample(b, score.copy(), 1.0)
            continue
        rb = GS.sampler._sample(b, score.copy(), 1.0)
        assert ra == rb and a.rng.rand() == b.rng.rand(), (case, ra, rb)
        assert np.array_equal(a.sub_score, b.sub_score, equal_nan=True), case
