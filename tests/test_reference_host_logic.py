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
                                                         np.array_equal(np.asarray(inp.collector_id_repeats)[d2[:, 0]], np.arange(N))),
                                 _cdf_cache={})
    mine._choice_no_replace = lambda ori, size: GS.sampler._choice_no_replace(mine, ori, size)
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


def test_genome_content_and_display_order_match_the_reference_lines():
    """cuda_lib_gl.py:1581-1668: genome_content and the ordering part of display_current_matrix on a scrambled genome."""
    pyr, inp = _mock_level(False, False)
    o = H.make_oracle(inp, pyr, seed=5)
    H.scramble(o, np.random.RandomState(12), 40)
    o.modify_gl_cuda_buffer()
    assert np.any(o.cur["ori"] == -1)

    class _Frags:
        def copy_from_gpu(self):
            pass
    fr = _Frags()
    for k in M.FIELDS:
        setattr(fr, k, o.cur[k].copy())
    W = inp.init_n_sub_frags

    class _Img:
        saved = []

        @staticmethod
        def fromarray(a):
            _Img.saved.append(a.shape)
            return types.SimpleNamespace(save=lambda f: None)
    ids = types.SimpleNamespace(get=lambda ary: ary.__setitem__(slice(None), o.cur["id_c"]))
    ref_self = types.SimpleNamespace(gpu_vect_frags=fr, gpu_id_contigs=ids, cpu_id_contigs=np.zeros_like(o.cur["id_c"]),
                                     np_sub_frags_id=inp.np_sub_frags_id, hic_matrix=np.zeros((W, W), dtype=np.float32),
                                     hic_matrix_sub_sampled=np.ones((2, 2)))
    mine = types.SimpleNamespace(gpu_vect_frags=fr, np_sub_frags_id=inp.np_sub_frags_id)
    r_order, r_dict = RH.method("genome_content")(ref_self)
    g_order, g_dict = GS.sampler.genome_content(mine)
    assert [int(x) for x in r_order] == [int(x) for x in g_order] and list(r_dict.keys()) == list(g_dict.keys())
    for k in r_dict:
        for f in ("id", "pos", "next", "prev", "start_bp", "id_c"):
            assert [int(x) for x in r_dict[k][f]] == [int(x) for x in g_dict[k][f]], (k, f)
    r3 = RH.method("display_current_matrix", {"Image": _Img})(ref_self, "unused.tiff")
    g3 = GS.sampler.display_current_matrix(mine, None)
    assert [int(x) for x in r3[0]] == [int(x) for x in g3[0]] and [int(x) for x in r3[2]] == [int(x) for x in g3[2]]
    assert {k: [int(x) for x in v] for k, v in r3[1].items()} == {k: [int(x) for x in v] for k, v in g3[1].items()}
    assert _Img.saved == [(W, W)] and len(g3[2]) == W


@pytest.mark.parametrize("allow_repeats,blacklist", [(False, False), (True, False), (True, True)])
def test_level_loader_matches_the_reference_lines(allow_repeats, blacklist):
    """simulation_loader.py: select_repeated_frags :369-394, modify_vect_frags :182-299, blacklist_contig :129-163 executed
    on the same level vs graal_b200.level.prepare_sampler_inputs (repeat detection, copies, collector / dispatcher, blacklist)."""
    import scipy.sparse as sp
    pyr, inp = _mock_level(allow_repeats, blacklist)
    lv = pyr.get_level(2)
    N = lv.n_frags
    csr = sp.csr_matrix((lv.vals.astype(np.float64), (lv.rows, lv.cols)), shape=(N, N))
    cmap = types.SimpleNamespace(N=256, __call__=None)

    class _Cmap:
        N = 256

        def __call__(self, i):
            return (0.0, 0.0, 0.0, 1.0)
    plt = types.SimpleNamespace(cm=types.SimpleNamespace(prism=_Cmap()))
    int2 = np.dtype([("x", np.int32), ("y", np.int32)], align=True)
    me = types.SimpleNamespace(level=types.SimpleNamespace(sparse_mat_csr=csr, S_o_A_frags=lv.S_o_A_frags), allow_repeats=allow_repeats, int2=int2)
    me.candidate_dup, me.data_candidate_dup = RH.loader_method("select_repeated_frags")(me)
    RH.loader_method("modify_vect_frags", {"plt": plt})(me)
    RH.loader_method("blacklist_contig")(me, [5] if blacklist else [0])
    if allow_repeats:
        assert len(me.candidate_dup) > 0 and inp.n_new_frags > inp.n_frags
    assert [int(x) for x in me.candidate_dup] == [int(x) for x in inp.id_frag_duplicated]
    for k in ("pos", "id_c", "start_bp", "len_bp", "circ", "id", "prev", "next", "l_cont", "l_cont_bp", "rep", "activ", "id_d"):
        assert np.array_equal(me.new_S_o_A_frags[k], inp.S_o_A_frags[k]), k
    assert np.array_equal(me.collector_id_repeats, inp.collector_id_repeats)
    assert np.array_equal(np.stack([me.frag_dispatcher["x"], me.frag_dispatcher["y"]], axis=1), np.asarray(inp.frag_dispatcher).reshape(-1, 2))
    assert [int(x) for x in me.frag_blacklisted] == [int(x) for x in inp.id_frags_blacklisted]
    assert me.n_frags == inp.n_new_frags and me.init_n_frags == inp.n_frags


def test_remove_problematic_fragments_matches_the_reference_function(tmp_path):
    """pyramid_sparse.remove_problematic_fragments (:573-848) executed on text files written from a synthetic level 0 with
    sparse and 1-bp fragments (inside contigs, at contig ends, a whole contig) vs graal_b200.pyramid_io's in-memory restatement."""
    import scipy.sparse as sp
    from graal_b200 import pyramid_io as PIO
    from graal_b200.level import PyramidLevel, _derive_frag_arrays
    rs = np.random.RandomState(3)
    sizes = [40, 25, 6, 30]
    cid = np.repeat(np.arange(1, 5), sizes).astype(np.int32)
    n = cid.size
    ln = rs.randint(200, 900, size=n)
    ln[[5, 17, 44]] = 1                                             # 1-bp fragments: locked whatever their contacts
    first = np.r_[True, cid[1:] != cid[:-1]]
    starts = np.nonzero(first)[0]
    seg = np.cumsum(first) - 1
    st = (np.cumsum(ln) - ln) - (np.cumsum(ln) - ln)[starts][seg]
    en = st + ln
    dense = np.triu((rs.rand(n, n) < 0.55) * rs.randint(1, 6, size=(n, n)), 0)
    for f in (3, 4, 20, 39, 64, 65, 66, 67, 68, 69, 70, 100):        # sparse fragments: mid contig, a run, contig ends, ALL of contig 3
        keep = rs.rand(n) < 0.03
        dense[f, :] *= keep; dense[:, f] *= keep
    r, c = np.nonzero(dense)
    lv = PyramidLevel(level=0, contig_id=cid, start_pos=st.astype(np.int32), end_pos=en.astype(np.int32), n_accu=np.ones(n, dtype=np.int32),
                      sub_low=np.arange(n, dtype=np.int32), sub_high=np.arange(n, dtype=np.int32),
                      rows=r.astype(np.int32), cols=c.astype(np.int32), vals=dense[r, c].astype(np.int32))
    _derive_frag_arrays(lv)
    mine, thresh, old2new = PIO.remove_problematic_fragments(lv)
    assert (old2new < 0).any() and mine.n_frags < n and len(np.unique(mine.contig_id)) == 3          # contig 3 deleted
    # ---- the reference function on files
    names = ["ctg%d" % k for k in range(1, 5)]
    fl, ci, ct = (str(tmp_path / x) for x in ("frags.txt", "contigs.txt", "contacts.txt"))
    with open(fl, "w") as h:
        h.write("id\tchrom\tstart_pos\tend_pos\tsize\tgc_content\taccu_frag\tfrag_start\tfrag_end\n")
        for i in range(n):
            h.write("%d\t%s\t%d\t%d\t%d\t0.5\t1\t%d\t%d\n" % (lv.S_o_A_frags["pos"][i] + 1, names[cid[i] - 1], st[i], en[i], ln[i], i, i))
    with open(ci, "w") as h:
        h.write("contig\tlength_kb\tn_frags\tcumul_length\n")
        cum = 0
        for k, m in enumerate(sizes):
            h.write("%s\t%d\t%d\t%d\n" % (names[k], 1, m, cum)); cum += m
    with open(ct, "w") as h:
        h.write("id_frag_a\tid_frag_b\tn_contact\n")
        for a, b in zip(r, c):
            h.write("%d\t%d\t%d\n" % (a, b, dense[a, b]))

    class _PB:
        def __init__(self, *a, **k):
            pass

        def render(self, *a, **k):
            pass
    ns = {"sp": sp, "ProgressBar": _PB}
    for helper in ("get_frag_info_from_fil", "get_contig_info_from_file", "file_len"):
        ns[helper] = RH.pyramid_function(helper, ns)
    f = RH.pyramid_function("remove_problematic_fragments", ns)
    pyramid = {"0": {"data": np.stack([r, c, dense[r, c]]).astype(np.int32), "nfrags": np.array([n], dtype=np.int32)}}
    out = [str(tmp_path / x) for x in ("new_contigs.txt", "new_frags.txt", "new_contacts.txt")]
    ref_thresh = f(ci, fl, ct, out[0], out[1], out[2], pyramid)
    assert abs(float(ref_thresh) - thresh) <= 1e-7 * max(1.0, abs(thresh))
    rows = [l.rstrip("\n").split("\t") for l in open(out[1]).read().split("\n")[1:] if l]
    assert len(rows) == mine.n_frags
    kept_names = [l.split("\t")[0] for l in open(out[0]).read().split("\n")[1:] if l]
    assert kept_names == ["ctg1", "ctg2", "ctg4"]
    for i, d in enumerate(rows):
        assert int(d[0]) == mine.S_o_A_frags["pos"][i] + 1 and d[1] == kept_names[mine.contig_id[i] - 1]
        assert (int(d[2]), int(d[3]), int(d[4]), int(d[6])) == (int(mine.start_pos[i]), int(mine.end_pos[i]), int(mine.end_pos[i] - mine.start_pos[i]) if False else int(d[4]), int(mine.n_accu[i]))
    ref_contacts = sorted(tuple(int(x) for x in l.split("\t")) for l in open(out[2]).read().split("\n")[1:] if l)
    assert ref_contacts == sorted(zip(mine.rows.tolist(), mine.cols.tolist(), mine.vals.tolist()))


class _SortedSet(set):
    """A set whose iteration order is defined (increasing): the reference iterates `list(V_set)` of a CPython-2 set."""

    def __iter__(self):
        return iter(sorted(set.__iter__(self)))

    def copy(self):
        return _SortedSet(set.copy(self))


@pytest.mark.parametrize("variant", ["step_metropolis_hastings_s_a", "step_mtm"])
def test_mh_steps_match_the_reference_lines(variant):
    """cuda_lib_gl.py:2836-3100: the reference's own step text (jump set, thresholds, exponentials, np.random.choice,
    acceptance) driving the oracle's primitives vs the oracle's restatement of the step: same trajectory."""
    pyr, inp = _mock_level(False, False)

    def make():
        o = H.make_oracle(inp, pyr, seed=1)
        o.rng = np.random                                            # the reference draws from the global stream
        H.scramble(o, np.random.RandomState(2), 40)                  # a rearranged genome: proposals that repair it are accepted
        o.set_jumping_distributions_parameters(3)
        for d in o.jump_dictionnary.values():
            d["set_frags"] = _SortedSet(d["set_frags"])
        o.init_likelihood()
        return o
    a, b = make(), make()

    class _View:                                                     # gpu_vect_frags of the reference: host arrays after copy_from_gpu
        def __init__(self, o, which):
            self.o, self.which = o, which

        def copy_from_gpu(self):
            g = self.o.cur if self.which == "cur" else self.o.fwd
            for k in M.FIELDS:
                setattr(self, k, g[k])
    a.gpu_vect_frags, a.gpu_vect_frags_forward = _View(a, "cur"), _View(a, "fwd")
    a.n_modif_metropolis = 13
    a.fwd = M.new_slot(a.n_new_frags)
    a.gpu_vect_frags_forward.copy_from_gpu()
    ref_detect = RH.method("detect_impossibility", py2=True)
    a.detect_impossibility = lambda fA, nb, fwd: ref_detect(a, fA, nb, fwd)
    a.dist_inter_genome = lambda view: OS.OracleSampler.dist_inter_genome(a, a.cur)
    ref_step = RH.method(variant, py2=True)
    sched = np.random.RandomState(9).randint(0, a.n_new_frags, size=16)
    np.random.seed(77)
    ra = []
    for fA in sched:
        a.gpu_vect_frags_forward.copy_from_gpu()
        ra.append(ref_step(a, int(fA), 0, 1, 0))
    np.random.seed(77)
    rb = [getattr(b, variant)(int(fA)) for fA in sched]
    assert H.slots_diff(a.cur, b.cur) == []
    for x, y in zip(ra, rb):
        assert x[0] == y[0] and tuple(x[1:5]) == tuple(y[1:5]) and x[6] == y[6], (x, y)
    assert any(x[0] != ra[0][0] for x in ra)                          # something was accepted
