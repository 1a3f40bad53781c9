"""The remaining entry points of the reference ``sampler`` (graal_b200/variants.py, oracle/variants.py) against the
reference's OWN Python lines (cuda_lib_gl.py) executed under Python 3 by tests/ref_host.py: the linear candidate draw of
debug_step_max_likelihood / step_max_likelihood_4_visu, define_neighbourhood + old_return_neighbours, local_flip (its kernel
calls served by oracle.mutations, themselves pinned to the compiled reference kernels), modify_genome, the parameter
packers.  Runs where /root/reference exists."""
import types

import numpy as np
import pytest

import ref_host as RH
import helpers as H
from graal_b200 import variants as GV
from graal_b200 import sampler as GS
from graal_b200.level import prepare_sampler_inputs, build_synthetic_pyramid
from oracle import mutations as M
from oracle import sampler as OS
from oracle import variants as OV

pytestmark = pytest.mark.skipif(not RH.available(), reason="reference sources not present")
NT = 13


@pytest.mark.parametrize("which", ["debug_step_max_likelihood", "step_max_likelihood_4_visu"])
def test_linear_draw_matches_the_reference_lines(which):
    """cuda_lib_gl.py:2228-2261 / :3242-3287: weights linear in the shifted scores."""
    debug = which.startswith("debug")
    last = ("sample_out = np.random.choice(id_ok_4_sampling[0], 1, p=sub_score)[0]" if debug
            else "sample_out = np.random.choice(id_ok_4_sampling[0], 1, p=self.sub_score)[0]")
    code = RH.block(which, "scores_2_remove = []", last)
    gen = np.random.RandomState(1)
    drawn = 0
    for case in range(300):
        n_nb = int(gen.randint(1, 5))
        score = -1e6 + gen.randn(NT * n_nb) * gen.choice([0.01, 1.0, 5.0, 40.0, 400.0])
        if debug:
            score = score.astype(np.float32)
        if case % 7 == 0:
            score[gen.randint(score.size)] += 1000.0
        F_t = float(gen.choice([1.0, 1.7, 0.6]))
        seed = int(gen.randint(1 << 30))
        ref_self = types.SimpleNamespace(score=score.copy(), n_tmp_struct=NT, temperature=lambda t, n: F_t, sub_score=None)
        ns = {"self": ref_self, "np": np, "time": __import__("time"), "t": 0, "n_step": 1}
        np.random.seed(seed)
        exec(code, ns)
        ref_out, ref_next = int(ns["sample_out"]), np.random.rand()
        for fn in (GV.linear_score_draw, OV.linear_score_draw):
            np.random.seed(seed)
            got, _ = fn(score.copy(), NT, 600 if debug else 30, None if debug else F_t, np.random, empty_is_max=not debug)
            assert got == ref_out and np.random.rand() == ref_next, (which, case, got, ref_out)
        drawn += int(ref_next != np.random.RandomState(seed).rand())
    assert drawn > 100


def _level(allow_repeats=False):
    pyr = build_synthetic_pyramid([300_000, 200_000, 150_000, 90_000, 40_000, 6_000], 600, 3, seed=11, cis_rowsum=300.0, v_inter=0.05)
    return pyr, prepare_sampler_inputs(pyr, 2, allow_repeats=allow_repeats)


@pytest.mark.parametrize("allow_repeats", [False, True])
def test_old_proposal_rule_matches_the_reference_lines(allow_repeats):
    """define_neighbourhood (:2548-2561) and old_return_neighbours (:2333-2360)."""
    pyr, inp = _level(allow_repeats)
    N, n = int(inp.n_frags), int(inp.n_new_frags)
    r, c, v = inp.level_coo
    dense = np.zeros((N, N), dtype=np.float32)
    dense[r, c] = v; dense[c, r] = v
    np.fill_diagonal(dense, 0)
    nv = np.asarray(inp.norm_vect_accu, dtype=np.float32)
    ref_self = types.SimpleNamespace(n_frags=N, hic_matrix_sub_sampled=dense, norm_vect_accu=np.matrix(nv.reshape(1, -1)) if nv.ndim < 2 else nv)
    RH.method("define_neighbourhood")(ref_self)
    mine = GV.sorted_neighbours_of((r, c, v), N, inp.norm_vect_accu)
    o = H.make_oracle(inp, pyr, seed=3)
    o.define_neighbourhood()
    for i in range(N):
        ref_line = np.asarray(ref_self.sorted_neighbours[i])
        vals = np.asarray(ref_self.matrix_normalized)[i]
        assert mine[i].size == ref_line.size == N - 1
        # same values along the order (which of two EQUAL counts comes first is the unstable argsort's choice) ...
        assert np.array_equal(vals[mine[i]], vals[ref_line]), i
        assert np.array_equal(mine[i], o.sorted_neighbours[i]), i
        # ... and the same bins wherever the order is defined
        strict = np.r_[True, np.diff(vals[ref_line]) > 0] & np.r_[np.diff(vals[ref_line]) > 0, True]
        assert np.array_equal(mine[i][strict], ref_line[strict]), i
    # old_return_neighbours with the reference's table on both sides
    disp = np.zeros(N, dtype=[("x", np.int32), ("y", np.int32)])
    d2 = np.asarray(inp.frag_dispatcher).reshape(-1, 2)
    disp["x"], disp["y"] = d2[:, 0], d2[:, 1]
    state = types.SimpleNamespace(id_d=np.asarray(inp.S_o_A_frags["id_d"]))
    ref2 = types.SimpleNamespace(gpu_vect_frags=state, sorted_neighbours=ref_self.sorted_neighbours,
                                 id_frag_duplicated=list(np.asarray(inp.id_frag_duplicated)), frag_dispatcher=disp,
                                 collector_id_repeats=np.asarray(inp.collector_id_repeats), id_frags_blacklisted=[3, 17])
    ref_fn = RH.method("old_return_neighbours")
    mine_self = types.SimpleNamespace(h_id_d=np.asarray(inp.S_o_A_frags["id_d"]), sorted_neighbours=[np.asarray(x) for x in ref_self.sorted_neighbours],
                                      _dup_set=set(int(f) for f in np.asarray(inp.id_frag_duplicated)), _black_set={3, 17},
                                      frag_dispatcher=d2, collector_id_repeats=np.asarray(inp.collector_id_repeats))
    o.sorted_neighbours = mine_self.sorted_neighbours
    o.id_frags_blacklisted = [3, 17]
    gen = np.random.RandomState(5)
    for case in range(200):
        fA, delta = int(gen.randint(n)), int(gen.choice([1, 3, 5]))
        a = [int(e) for e in ref_fn(ref2, fA, delta)]
        assert a == GV.VariantsMixin.old_return_neighbours(mine_self, fA, delta), (case, fA)
        assert a == o.old_return_neighbours(fA, delta), (case, fA)
    if allow_repeats:
        assert len(mine_self._dup_set) > 0


class _Slot(dict):
    """An oracle slot that also answers the reference's GPUStruct calls (get_ptr / copy_from_gpu / attribute access)."""
    def get_ptr(self):
        return self

    def copy_from_gpu(self):
        pass

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def _kernel_mocks(ref_self):
    """The reference's kernel handles (called with block= / grid= keywords) served by oracle.mutations."""
    def strip(f):
        return lambda *a, **kw: f(*a)
    ref_self.simple_copy = strip(lambda dst, src, n: M.simple_copy(dst, src))
    ref_self.pop_out = strip(lambda dst, src, ids, f, max_id, n: M.pop_out_frag(dst, src, ids, int(f), max_id))
    ref_self.pop_in_3 = strip(lambda dst, src, fp, fi, max_id, ori, n: M.pop_in_frag_3(dst, src, int(fp), int(fi), max_id, int(ori)))
    ref_self.pop_in_4 = strip(lambda dst, src, fp, fi, max_id, ori, n: M.pop_in_frag_4(dst, src, int(fp), int(fi), max_id, int(ori)))
    ref_self.flip_frag = strip(lambda dst, src, f, n: M.flip_frag(dst, src, int(f)))


class _Event:
    def record(self, *a):
        pass

    def synchronize(self):
        pass


def test_local_flip_matches_the_reference_lines():
    """cuda_lib_gl.py:1056-1154 with its kernels served by oracle.mutations vs oracle.variants.local_flip; and what the move
    means: the window around fA comes back in reverse order with every orientation flipped."""
    pyr, inp = _level()
    o = H.make_oracle(inp, pyr, seed=2)
    rng = np.random.RandomState(4)
    H.scramble(o, rng, 25)
    o.modify_gl_cuda_buffer()
    n = o.n_new_frags
    fn = RH.method("local_flip", extra_ns={"cuda": types.SimpleNamespace(Event=_Event)}, py2=True)
    checked = 0
    for case in range(40):
        fA = int(rng.randint(n))
        mode = int(rng.choice([12, 13, 14, 15]))
        max_id = np.int32(o.cur["id_c"].max())
        ref_self = types.SimpleNamespace(n_new_frags=n, gpu_vect_frags=_Slot(M.copy_slot(o.cur)), scrambled_gpu_vect_frags=_Slot(M.new_slot(n)),
                                         pop_gpu_vect_frags=_Slot(M.new_slot(n)), pop_gpu_id_contigs=np.zeros(n, dtype=np.int32),
                                         collector_gpu_vect_frags={mode: _Slot(M.new_slot(n))})
        _kernel_mocks(ref_self)
        fn(ref_self, fA, mode, max_id)
        ws = M.Workspace(n)
        scr, col = M.new_slot(n), M.new_slot(n)
        OV.local_flip(ws, o.cur, scr, col, fA, mode, max_id)
        assert H.slots_diff(dict(ref_self.collector_gpu_vect_frags[mode]), col) == [], (case, fA, mode)
        # the meaning of the move, wherever the window holds fA's neighbours on both sides inside a linear contig
        c = o.cur
        if c["circ"][fA] == 0 and M.check_invariants(col) == []:
            members = np.nonzero(c["id_c"] == c["id_c"][fA])[0]
            before = members[np.argsort(c["pos"][members])]
            d = mode - 11
            lo, hi = max(int(c["pos"][fA]) - d, 0), min(int(c["pos"][fA]) + d, len(before) - 1)
            after_members = np.nonzero(col["id_c"] == col["id_c"][fA])[0]
            after = after_members[np.argsort(col["pos"][after_members])]
            if len(after) == len(before):
                exp = np.concatenate([before[:lo], before[lo:hi + 1][::-1], before[hi + 1:]])
                assert np.array_equal(after, exp), (case, fA, mode)
                assert np.array_equal(col["ori"][before[lo:hi + 1]], -c["ori"][before[lo:hi + 1]])
                checked += 1
    assert checked > 10


def test_modify_genome_and_packers_match_the_reference_lines():
    """modify_genome (:1521-1537: same draws from the global stream, same committed mutations), modify_param_simu
    (:3131-3138), setup_rippe_parameters_4_simu (:1186-1201), setup_model_parameters (:1216-1227)."""
    pyr, inp = _level()
    o = H.make_oracle(inp, pyr, seed=2)
    n = o.n_new_frags
    # modify_genome: the reference text with test_copy_struct served by the oracle
    ref_o = H.make_oracle(inp, pyr, seed=2)
    cur = _Slot(ref_o.cur)
    ref_self = types.SimpleNamespace(n_new_frags=n, n_tmp_struct=NT, gpu_vect_frags=cur)
    ref_self.test_copy_struct = lambda fA, fB, mode, max_id: M.apply_mutation(ref_o.ws, ref_o.cur, int(fA), int(fB), int(mode), max_id, ref_o.id_contigs)
    fn = RH.method("modify_genome", extra_ns={"raw_input": lambda *_: (_ for _ in ()).throw(AssertionError("invariant broken"))}, py2=True)
    np.random.seed(77)
    fn(ref_self, 12)
    nxt = np.random.rand()
    o.rng = np.random
    np.random.seed(77)
    o.modify_genome(12)
    assert np.random.rand() == nxt and H.slots_diff(ref_o.cur, o.cur) == []
    # packers
    dt_rippe = GS.PARAM_DTYPE
    dt_exp = np.dtype([(k, np.float32) for k in GV.PARAM_SIMU_EXP_FIELDS], align=True)
    ref_p = types.SimpleNamespace(param_simu_T=dt_rippe, param_simu_exp=dt_exp, mean_value_trans=np.float32(0.031))
    mine = types.SimpleNamespace(mean_value_trans=np.float32(0.031))
    a = RH.method("setup_rippe_parameters_4_simu")(ref_p, 1.0, 9.6, -1.5, 3.0, 0.02, 850.0)
    b = GV.VariantsMixin.setup_rippe_parameters_4_simu(mine, 1.0, 9.6, -1.5, 3.0, 0.02, 850.0)
    # the reference ran under NumPy 1.x (float32 scalar op Python float -> float64); its text executed under NumPy 2 keeps
    # float32 there: c1 and fact may differ in the last bit, everything else is identical
    for k in a.dtype.names:
        if k in ("c1", "fact"):
            assert abs(float(a[k][0]) - float(b[k][0])) <= 2 * np.spacing(np.float32(abs(a[k][0]))), k
        else:
            assert a[k][0] == b[k][0], k
    a = RH.method("setup_model_parameters")(ref_p, (10.0, 200.0, -0.5, -1.2, -2.0, 33.0), 900.0)
    b = GV.VariantsMixin.setup_model_parameters(mine, (10.0, 200.0, -0.5, -1.2, -2.0, 33.0), 900.0)
    assert a.tobytes() == b.tobytes()
    p0 = GS.sampler.setup_rippe_parameters(types.SimpleNamespace(mean_value_trans=np.float32(0.031)), (1.0, 9.6, -1.5, 3.0, 700.0), 850.0) \
        if hasattr(GS.sampler, "setup_rippe_parameters") else b
    for id_val in (0, 1, 2):
        a = RH.method("modify_param_simu")(None, p0, id_val, 2.5)
        b = GV.VariantsMixin.modify_param_simu(None, p0, id_val, 2.5)
        assert a.tobytes() == b.tobytes()


def test_update_neighbourhood_matches_the_reference_lines():
    """cuda_lib_gl.py:717-733 on a matrix without ties (the dense argsort is only defined there)."""
    gen = np.random.RandomState(2)
    N = 40
    up = np.triu(gen.permutation(N * N).reshape(N, N).astype(np.float32) + 1, 1)
    dense = up + up.T
    r, c = np.nonzero(np.triu(dense, 1))
    ref_self = types.SimpleNamespace(hic_matrix_sub_sampled=dense, list_frag_to_sample=[0, 5, 7, 31], list_to_pop_out=[2, 9])
    RH.method("update_neighbourhood", py2=True)(ref_self)
    mine = types.SimpleNamespace(_level_coo=(r, c, dense[r, c]), n_frags=N, list_frag_to_sample=[0, 5, 7, 31], list_to_pop_out=[2, 9])
    GV.VariantsMixin.update_neighbourhood(mine)
    assert len(mine.sorted_neighbours) == 4
    for a, b in zip(ref_self.sorted_neighbours, mine.sorted_neighbours):
        assert np.array_equal(np.asarray(a), b)
