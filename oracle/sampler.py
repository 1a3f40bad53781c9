"""NumPy restatement of the reference sampler's host logic (TEST INFRASTRUCTURE).

Follows /root/reference/cuda_lib_gl.py (class sampler):
  __init__ data prep            :153-172 (zero diagonals, blacklist rows), :226-233, :444-446
  define_repeats                :452-469
  dist_inter_genome             :475-541
  eval_likelihood               :543-631
  estimate_parameters           :1229-1294  (+ optim_rippe_curve_update.py:53-61,73-135)
  explode_genome                :1539-1557
  modify_gl_cuda_buffer         :1695-1788  (relabel part)
  step_max_likelihood           :1793-1980
  step_nuisance_parameters      :2022-2107
  return_neighbours             :2295-2331
  setup_distri_frags            :2363-2390
  stream_likelihood             :2392-2546
The reference never seeds NumPy: here every draw comes from one explicit
``np.random.RandomState`` consumed in the reference's call order.
Works on the DENSE sub-level matrix, exactly like the reference: small levels only.
"""
import numpy as np
from scipy.optimize import leastsq, fsolve

from . import mutations as M
from . import likelihood as L
from .variants import OracleVariants

F32 = np.float32
I32 = np.int32
D_RIPPE = 3          # module constant ``d`` of optim_rippe_curve_update.py:9


# ---------------------------------------------------------------- optim_rippe_curve_update.py
def peval(x, param):
    """optim_rippe_curve_update.py:22-28 (param[3] is the amplitude, d is the module constant)."""
    d = D_RIPPE
    return param[3] * (0.53 * (param[0] ** -3.) * np.power((param[1] * x / param[0]), (param[2])) *
                       np.exp((d - 2) / ((np.power((param[1] * x / param[0]), 2) + d))))


def _log_residuals(p, y, x):
    kuhn, lm, slope, A = p
    d = D_RIPPE
    with np.errstate(all="ignore"):
        r = np.log(A) + np.log(0.53) - 3 * np.log(kuhn) + slope * (np.log(lm * x) - np.log(kuhn)) + \
            (d - 2) / ((np.power((lm * x / kuhn), 2) + d))
    return y - r


def estimate_param_rippe(y_meas, x_bins):
    """optim_rippe_curve_update.py:73-115 (incl. the slope check on the INITIAL constant, Q10)."""
    kuhn, lm, slope = 1, 9.6, -1.5
    A = np.sum(y_meas)
    p0 = [kuhn, lm, slope, A]
    plsq = leastsq(_log_residuals, p0, args=(np.log(y_meas), x_bins))
    y_estim = peval(x_bins, plsq[0])
    kuhn_x, lm_x, slope_x, A_x = plsq[0]
    out = [kuhn_x, lm_x, slope_x, D_RIPPE, A_x]
    if np.any(np.isnan(np.array(out))) or slope >= 0:
        out = [kuhn, lm, slope, D_RIPPE, A]
    return out, y_estim


def estimate_max_dist_intra(p, val_inter):
    """optim_rippe_curve_update.py:117-135."""
    kuhn, lm, slope, d, A = p

    def resid(x, q):
        kuhn, lm, slope, d, A, y = q
        with np.errstate(all="ignore"):
            r = A * (0.53 * (kuhn ** -3.) * np.power((lm * x / kuhn), slope) *
                     np.exp((d - 2) / ((np.power((lm * x / kuhn), 2) + d))))
        return y - r
    x = fsolve(resid, 500, args=([kuhn, lm, slope, d, A, val_inter]))
    return x[0]


def distance_histogram(sub_soa, hic_matrix, max_dist_kb, size_bin_kb):
    """The O(W^2) loop of estimate_parameters (cuda_lib_gl.py:1236-1270), vectorised per row:
    mean contacts (zeros included) of cis sub-frag pairs per distance bin; empty or 0 -> 1e-10."""
    bins = np.arange(size_bin_kb, max_dist_kb + size_bin_kb, size_bin_kb)
    sums = np.zeros(len(bins), dtype=np.float64)
    cnts = np.zeros(len(bins), dtype=np.int64)
    idc, st, ln, pos = (np.asarray(sub_soa[k]) for k in ("id_c", "start_bp", "len_bp", "pos"))
    W = idc.shape[0]
    for i in range(W - 1):
        j = np.arange(i + 1, W)
        j = j[idc[j] == idc[i]]
        if j.size == 0:
            continue
        fwd = pos[i] < pos[j]
        d = np.where(fwd, ((st[j] - st[i] - ln[i]) + (ln[i] + ln[j]) / 2.) / 1000.,
                     ((st[i] - st[j] - ln[j]) + (ln[j] + ln[i]) / 2.) / 1000.)
        ok = d < max_dist_kb
        b = (d[ok] / size_bin_kb).astype(np.int64)
        np.add.at(sums, b, hic_matrix[i, j[ok]].astype(np.float64))
        np.add.at(cnts, b, 1)
    mean = np.full(len(bins), 1e-10, dtype=F32)
    ok = (cnts > 0) & (sums > 0)
    mean[ok] = (sums[ok] / cnts[ok]).astype(F32)
    return bins, mean, sums, cnts


# ---------------------------------------------------------------- the sampler
class OracleSampler(OracleVariants):
    N_TMP = 13

    def __init__(self, inp, rng, hic_matrix=None, hic_matrix_sub_sampled=None):
        """``inp``: the constructor inputs of the reference sampler (any object with the attribute
        names of cuda_lib_gl.py:33-42).  ``rng``: np.random.RandomState."""
        self.rng = rng
        self.inp = inp
        self.n_frags = int(inp.n_frags)
        self.n_new_frags = int(inp.n_new_frags)
        self.id_frags_blacklisted = list(inp.id_frags_blacklisted)
        self.id_frag_duplicated = np.asarray(inp.id_frag_duplicated, dtype=I32)
        self.uniq_frags = np.setdiff1d(np.arange(self.n_frags, dtype=I32), self.id_frag_duplicated).astype(I32)
        self.collector = np.asarray(inp.collector_id_repeats, dtype=I32)
        self.dispatcher = np.asarray(inp.frag_dispatcher, dtype=I32).reshape(-1, 2)
        self.mean_value_trans = inp.mean_value_trans
        sub = inp.dense_sub_matrix() if hic_matrix is None else np.array(hic_matrix, dtype=F32)
        lvl = inp.dense_level_matrix() if hic_matrix_sub_sampled is None else np.array(hic_matrix_sub_sampled, dtype=F32)
        np.fill_diagonal(sub, 0)                                   # :157-160
        np.fill_diagonal(lvl, 0)
        for f in self.id_frags_blacklisted:                        # :161-172
            real = inp.S_o_A_frags["id_d"][f]
            lvl[real, :] = 0
            lvl[:, real] = 0
            da = inp.np_sub_frags_id[real]
            for k in range(da[3]):
                sub[da[k], :] = self.mean_value_trans
                sub[:, da[k]] = self.mean_value_trans
        self.hic_matrix = sub
        self.hic_matrix_sub_sampled = lvl
        self.lv = L.DenseLevel(self.n_frags, inp.np_sub_frags_id, inp.np_sub_frags_len_bp, inp.np_sub_frags_accu,
                               self.collector, self.dispatcher, sub, inp.mean_squared_frags_per_bin)
        s0 = inp.S_o_A_frags
        self.cur = {k: np.array(s0[k], dtype=I32) for k in M.FIELDS}
        self.cur["ori"][:] = 1                                     # :244,259 (Q5)
        self.ws = M.Workspace(self.n_new_frags)
        self.id_contigs = self.cur["id_c"].copy()
        self.sub_index = np.zeros(self.n_new_frags, dtype=I32)     # persistent gpu_sub_index (:219)
        self.np_init_prev = s0["prev"].copy()
        self.np_init_next = s0["next"].copy()
        self.np_init_ori = np.ones(self.n_new_frags, dtype=I32)
        self.np_init_orientable = (np.asarray(inp.np_sub_frags_id)[s0["id_d"], 3] > 1).astype(I32)
        self.n_neighbors = 10
        self.setup_distri_frags()
        self.define_repeats()
        self.param_simu = None
        self.likelihood_t = None
        self.curr_likelihood = None
        self.o = 0

    # ---- :452-469
    def define_repeats(self):
        c = self.cur
        rep_ids = np.unique(c["id_d"][c["id"] != c["id_d"]])
        self.is_repeat = np.isin(c["id_d"], rep_ids)
        tmp = list(self.id_frags_blacklisted) + list(np.nonzero(self.is_repeat)[0])
        self.n_frags_4_dist = len(np.unique(tmp))

    # ---- :2363-2390
    def setup_distri_frags(self):
        """top-10 of each LEVEL-matrix row by reversed argsort; ties are DEFINED by a stable sort
        (so, after the reversal, equal values come in decreasing column order), SURVEY H7."""
        self.distri_frags = {}
        for i in range(self.n_frags):
            v = self.hic_matrix_sub_sampled[i, :].astype(F32)
            xk = np.argsort(v, kind="stable")[::-1][: self.n_neighbors].astype(I32)
            dat = v[xk] ** 3
            if dat.sum() > 0:
                pk = dat / dat.sum()
            else:
                tmp = np.ones_like(dat, dtype=F32)
                pk = tmp / tmp.sum()
            self.distri_frags[i] = dict(xk=xk, pk=pk)

    # ---- :2295-2331
    def return_neighbours(self, id_fA, delta0):
        ori_id = self.cur["id_d"][id_fA]
        delta = min(self.n_neighbors, delta0)
        distri = self.distri_frags[ori_id]["pk"]
        n_max = min(delta, np.nonzero(distri != 0)[0].shape[0])
        init_id = self.rng.choice(self.distri_frags[ori_id]["xk"], n_max, p=distri, replace=False)
        out = []
        if ori_id in self.id_frag_duplicated:
            d = self.dispatcher[ori_id]
            out.extend(np.setdiff1d(self.collector[d[0]:d[1]], id_fA))
        for id_fB in init_id:
            d = self.dispatcher[id_fB]
            out.extend(self.collector[d[0]:d[1]])
        return [int(e) for e in out if e not in self.id_frags_blacklisted]

    # ---- :1203-1214, 1229-1294
    def set_params(self, kuhn, lm, slope, d, fact, d_max, v_inter=None):
        self.param_simu = L.make_params(kuhn, lm, slope, d, fact, d_max,
                                        self.mean_value_trans if v_inter is None else v_inter)

    def estimate_parameters(self, max_dist_kb, size_bin_kb):
        self.bins, self.mean_contacts, _, _ = distance_histogram(self.inp.S_o_A_sub_frags, self.hic_matrix,
                                                                 max_dist_kb, size_bin_kb)
        p, self.y_estim = estimate_param_rippe(self.mean_contacts, self.bins)
        d_max = estimate_max_dist_intra(p, self.mean_value_trans)
        self.set_params(p[0], p[1], p[2], p[3], p[4], d_max)

    # ---- :543-631
    def eval_likelihood(self, params=None):
        self.curr_likelihood = L.evaluate_likelihood(self.cur, self.lv, self.param_simu if params is None else params)
        return np.float64(self.curr_likelihood.sum())

    def init_likelihood(self):
        self.likelihood_t = self.eval_likelihood()

    # ---- :1695-1788 (relabel part)
    def modify_gl_cuda_buffer(self, id_fi=0, dt=0):
        max_id = M.relabel_contigs(self.cur)
        self.id_contigs[:] = self.cur["id_c"]
        return max_id

    # ---- :1539-1557
    def explode_genome(self, dt=0):
        for i in range(self.n_new_frags):
            self.modify_gl_cuda_buffer(i, dt)
            max_id = I32(self.cur["id_c"].max())
            M.apply_mutation(self.ws, self.cur, i, 0, 0, max_id, self.id_contigs)

    # ---- :1559-1578
    def apply_replay_simu(self, id_fA, id_fB, op_sampled, dt=0):
        self.modify_gl_cuda_buffer(id_fA, dt)
        max_id = I32(self.cur["id_c"].max())
        M.apply_mutation(self.ws, self.cur, id_fA, id_fB, op_sampled, max_id, self.id_contigs)

    # ---- :2392-2546
    def candidate_index_sets(self, id_fA, id_fB):
        """fill_sub_index_fA/fB (kernels3.cu:3225-3249) on the persistent sub_index, then the numpy
        set operations of stream_likelihood (cuda_lib_gl.py:2441-2470)."""
        c = self.cur
        contig_A, len_A = c["id_c"][id_fA], int(c["l_cont"][id_fA])
        inA = c["id_c"] == contig_A
        self.sub_index[c["pos"][inA]] = c["id_d"][inA]
        contig_B, len_B = c["id_c"][id_fB], int(c["l_cont"][id_fB])
        if contig_B != contig_A:
            inB = c["id_c"] == contig_B
            self.sub_index[len_A + c["pos"][inB]] = c["id_d"][inB]
            size = len_A + len_B
        else:
            size = len_A
        init = self.sub_index[:size]
        return np.setdiff1d(init, self.id_frag_duplicated), np.intersect1d(init, self.id_frag_duplicated)

    def stream_likelihood(self, id_fA, id_fB, id_x, likelihood_t, max_id):
        M.perform_modifications(self.ws, self.cur, id_fA, id_fB, max_id)
        no_rep, rep = self.candidate_index_sets(id_fA, id_fB)
        for j in range(self.N_TMP):
            d = L.sub_compute_likelihood(self.ws.collector[j], self.lv, self.param_simu, self.curr_likelihood,
                                         no_rep, rep, self.uniq_frags)
            self.delta[id_x * self.N_TMP + j] = d
            self.score[id_x * self.N_TMP + j] = d + likelihood_t

    # ---- :1793-1980
    def step_max_likelihood(self, id_fA, delta, size_block=512, dt=0, t=0, n_step=1):
        c = self.cur
        if id_fA not in self.id_frags_blacklisted:
            n_contigs = len(np.unique(c["id_c"]))
            max_id = self.modify_gl_cuda_buffer(id_fA, dt)
            id_start = np.nonzero(c["start_bp"] == 0)[0]
            mean_len_bp = c["l_cont_bp"][id_start].mean()
            max_len, min_len = c["l_cont"].max(), c["l_cont"].min()
            likelihood_t = self.eval_likelihood()
            self.likelihood_t = likelihood_t
            max_id = I32(self.id_contigs.max())
            id_neighbours = self.return_neighbours(id_fA, delta)
            n_nb = len(id_neighbours)
            self.score = np.zeros(n_nb * self.N_TMP, dtype=np.float64)
            self.delta = np.zeros(n_nb * self.N_TMP, dtype=np.float64)
            id_neighbours.sort()
            self.id_neighbours = id_neighbours
            for id_x in range(n_nb):
                self.stream_likelihood(id_fA, id_neighbours[id_x], id_x, likelihood_t, max_id)
            sample_out, self.sample_margin = sample_candidate(self.score, self.N_TMP, self.rng, self.temperature(t, n_step))
            id_f_sampled = id_neighbours[sample_out // self.N_TMP]
            op_sampled = sample_out % self.N_TMP
            M.apply_mutation(self.ws, c, id_fA, id_f_sampled, op_sampled, max_id, self.id_contigs)
            o = self.score[sample_out]
            self.o = o
        else:
            o = self.o
            n_contigs = len(np.unique(c["id_c"]))
            self.modify_gl_cuda_buffer(id_fA, dt)
            id_start = np.nonzero(c["start_bp"] == 0)[0]
            mean_len_bp = c["l_cont_bp"][id_start].mean()
            max_len, min_len = c["l_cont"].max(), c["l_cont"].min()
            op_sampled, id_f_sampled = -1, id_fA
        F_t = self.temperature(t, n_step)
        dist = self.dist_inter_genome(c)
        self.likelihood_t = o
        return o, n_contigs, min_len, mean_len_bp, max_len, op_sampled, id_f_sampled, dist, F_t

    def temperature(self, t, n_step):
        return 1.0

    # ---- :2022-2107
    def step_nuisance_parameters(self, dt=0, t=0, n_step=1):
        p = self.param_simu
        kuhn, lm, c1, slope, d, d_max, fact, d_nuc = (p[k] for k in L.PARAM_FIELDS)
        sigma_fact = 10 ** (np.log10(fact) - 2)
        id_modif = self.rng.choice(4)
        if id_modif == 0:
            new_fact = np.float64(fact) + self.rng.normal(loc=0.0, scale=sigma_fact)   # NumPy 1.x: float32 + float -> float64
            new_d_max = estimate_max_dist_intra([kuhn, lm, slope, d, new_fact], d_nuc)
            out = (kuhn, lm, slope, d, new_fact, new_d_max, d_nuc)
        elif id_modif == 1:
            new_slope = np.float64(slope) + self.rng.normal(loc=0.0, scale=0.05)
            new_d_max = estimate_max_dist_intra([kuhn, lm, new_slope, d, fact], d_nuc)
            out = (kuhn, lm, new_slope, d, fact, new_d_max, d_nuc)
        elif id_modif == 2:
            new_d_max = np.float64(d_max) + self.rng.normal(loc=0.0, scale=100)
            new_d_nuc = peval(new_d_max, [kuhn, lm, slope, d, fact])       # Q10: param[3] = d is the amplitude
            out = (kuhn, lm, slope, d, fact, new_d_max, new_d_nuc)
        else:
            new_d_nuc = np.float64(d_nuc) + self.rng.normal(loc=0.0, scale=0.5)
            new_d_max = estimate_max_dist_intra([kuhn, lm, slope, d, fact], new_d_nuc)
            out = (kuhn, lm, slope, d, fact, new_d_max, new_d_nuc)
        k_, lm_, sl_, d_, f_, dm_, dn_ = out
        test = L.make_params(k_, lm_, sl_, d_, f_, dm_, dn_)
        test_likelihood = self.eval_likelihood(test)              # overwrites curr_likelihood (Q11)
        F_t = self.temperature(t, n_step)
        with np.errstate(over="ignore"):
            ratio = np.exp((test_likelihood - self.likelihood_t) / F_t)
        u = self.rng.rand()
        success = 0
        if ratio >= u:
            success = 1
            self.param_simu = test
            self.likelihood_t = test_likelihood
        p = self.param_simu
        y_rippe = peval(self.bins, [p["kuhn"], p["lm"], p["slope"], p["d"], p["fact"]]) if hasattr(self, "bins") else None
        return p["fact"], p["d"], p["d_max"], p["v_inter"], p["slope"], self.likelihood_t, success, y_rippe

    # ------------------------------------------------------------------------------------------------
    # Metropolis-Hastings / multiple-try variants (cuda_lib_gl.py:735-839, 957-1013, 2548-2588, 2615-3129); unused by the
    # shipped GUI path.  `list(V_set)` of a CPython-2 set has no defined order: increasing ids here and on the device.
    # ------------------------------------------------------------------------------------------------
    N_MH = 13

    def set_jumping_distributions_parameters(self, delta):
        """:2548-2588 on the dense level matrix (stable argsort where the reference's is unstable on ties)."""
        nv = np.asarray(self.inp.norm_vect_accu, dtype=np.float64).reshape(1, -1)
        mat_norm = np.array(nv.T * nv, dtype=F32)
        self.matrix_normalized = self.hic_matrix_sub_sampled / mat_norm
        tmp_sorted = self.matrix_normalized.argsort(axis=1, kind="stable")
        self.jump_dictionnary = dict()
        for i in range(self.n_frags):
            line = [int(x) for x in tmp_sorted[i, :] if x != i]
            ids = np.array(line[-delta:], dtype=I32)
            scores = np.array(self.matrix_normalized[i, ids], dtype=F32)
            with np.errstate(all="ignore"):
                norm_scores = scores / scores.sum()
            self.jump_dictionnary[i] = {"proba": norm_scores, "frags": ids, "set_frags": set(int(x) for x in ids)}

    def _mh_base(self, forward):
        if forward:
            return self.cur
        return self.fwd

    def all_modifications_metropolis(self, id_fA, id_fB, max_id, forward):
        """:2651-2657 with pop_out_pop_in_4_mh :735-789, split_4_mh :791-811, paste_4_mh :813-839, transloc_4_mh :957-1013."""
        ws, base = self.ws, self._mh_base(forward)
        for mode in range(6):
            M.pop_out_frag(ws.pop, base, ws.pop_id_contigs, id_fA, max_id)
            max_id2 = I32(ws.pop_id_contigs.max())
            dst = ws.collector[mode]
            if mode == 0:
                M.simple_copy(dst, ws.pop)
            elif mode == 1:
                M.flip_frag(dst, base, id_fA)
            elif mode in (2, 3):
                M.pop_in_frag_3(dst, ws.pop, id_fA, id_fB, max_id2, 1 if mode == 2 else -1)
            else:
                M.pop_in_frag_4(dst, ws.pop, id_fA, id_fB, max_id2, 1 if mode == 4 else -1)
        for up in (0, 1):
            M.split_contig(ws.collector[6 + up], base, ws.trans1_id_contigs, id_fA, up, max_id)
        ext = lambda f: base["prev"][f] == -1 or base["next"][f] == -1
        if ext(id_fA) and ext(id_fB):
            M.paste_contigs(ws.collector[8], base, id_fA, id_fB, max_id)
        else:
            M.simple_copy(ws.collector[8], base)
        mode = 0
        for up_a in (0, 1):
            M.split_contig(ws.trans1, base, ws.trans1_id_contigs, id_fA, up_a, max_id)
            for up_b in (0, 1):
                max_id1 = I32(ws.trans1_id_contigs.max())
                ok = (base["next"][id_fB] == -1) if up_b == 0 else (base["prev"][id_fB] == -1)
                dst = ws.collector[9 + mode]
                if ok:
                    M.split_contig(ws.trans2, ws.trans1, ws.trans2_id_contigs, id_fB, up_b, max_id1)
                    M.paste_contigs(dst, ws.trans2, id_fA, id_fB, I32(ws.trans2_id_contigs.max()))
                else:
                    M.simple_copy(dst, base)
                mode += 1

    def compute_all_score_MH(self, id_fA, V_set, forward):
        """:2615-2649 + multi_likelihood_4_metropolis :2659-2806."""
        base = self._mh_base(forward)
        list_fB = sorted(int(x) for x in V_set)
        vect = L.evaluate_likelihood(base, self.lv, self.param_simu)
        if forward:
            self.curr_likelihood = vect
        else:
            self.curr_likelihood_forward = vect
        likelihood_t = np.float64(vect.sum())
        max_id = I32(base["id_c"].max())
        score = np.zeros(self.N_MH * len(list_fB), dtype=np.float64)
        contig_A, len_A = base["id_c"][id_fA], int(base["l_cont"][id_fA])
        inA = base["id_c"] == contig_A
        self.sub_index[base["pos"][inA]] = base["id_d"][inA]
        for x, id_fB in enumerate(list_fB):
            self.all_modifications_metropolis(id_fA, id_fB, max_id, forward)
            contig_B, len_B = base["id_c"][id_fB], int(base["l_cont"][id_fB])
            size = len_A
            if contig_B != contig_A:
                inB = base["id_c"] == contig_B
                self.sub_index[len_A + base["pos"][inB]] = base["id_d"][inB]
                size = len_A + len_B
            init = self.sub_index[:size]
            no_rep, rep = np.setdiff1d(init, self.id_frag_duplicated), np.intersect1d(init, self.id_frag_duplicated)
            for j in range(self.N_MH):
                score[x * self.N_MH + j] = L.sub_compute_likelihood(self.ws.collector[j], self.lv, self.param_simu, vect,
                                                                    no_rep, rep, self.uniq_frags) + likelihood_t
        return score

    def udpate_forward_vect(self, id_fA, id_fB, id_op, max_id):
        """:2808-2834."""
        self.all_modifications_metropolis(id_fA, id_fB, max_id, True)
        if not hasattr(self, "fwd") or self.fwd is None:
            self.fwd = M.new_slot(self.n_new_frags)
        M.simple_copy(self.fwd, self.ws.collector[int(id_op)])

    def validate_struct(self, id_fA, id_f_sampled, id_op, max_id):
        """:3102-3129."""
        self.all_modifications_metropolis(id_fA, id_f_sampled, max_id, True)
        M.copy_struct(self.cur, self.ws.collector[int(id_op)], self.id_contigs)
        self.init_likelihood()

    def detect_impossibility(self, id_fA, list_neighbours, forward):
        """:3072-3100."""
        g = self._mh_base(forward)
        out = []
        fA_ok = g["prev"][id_fA] == -1 or g["next"][id_fA] == -1
        for idx, fB in enumerate(list_neighbours):
            fB_ok = g["prev"][fB] == -1 or g["next"][fB] == -1
            if not (fB_ok and fA_ok):
                out.append(self.N_MH * idx + 8)
            if not g["next"][fB] == -1:
                out += [self.N_MH * idx + 9, self.N_MH * idx + 11]
            if not g["prev"][fB] == -1:
                out += [self.N_MH * idx + 10, self.N_MH * idx + 12]
        return out

    def _mh_prologue(self, id_fA, dt):
        c = self.cur
        stats = (len(np.unique(c["id_c"])), c["l_cont"].min(), c["l_cont"].mean(), c["l_cont"].max())
        max_id = self.modify_gl_cuda_buffer(id_fA, dt)
        V_set = set(self.jump_dictionnary[id_fA]["set_frags"])
        if c["prev"][id_fA] != -1:
            V_set.add(int(c["prev"][id_fA]))
        if c["next"][id_fA] != -1:
            V_set.add(int(c["next"][id_fA]))
        return stats, max_id, V_set, np.array(sorted(V_set), dtype=I32)

    def _mh_accept(self, ratio, id_fA, f_star, omega_star, max_id, star):
        r = np.min([1, ratio])
        if r == 1 or r >= self.rng.rand():
            self.validate_struct(id_fA, f_star, omega_star, max_id)
            self.likelihood_t = star

    def step_metropolis_hastings_s_a(self, id_fA, t=0, n_step=1, dt=0):
        """:2836-2934."""
        (n_contigs, min_len, mean_len, max_len), max_id, V_set, nb = self._mh_prologue(id_fA, dt)
        F_t = self.temperature(t, n_step)
        lf = self.compute_all_score_MH(id_fA, V_set, True)
        s = lf / F_t
        mx = s.max()
        s[s <= mx - 10] = mx - 10
        s = s - s.min()
        sf = np.exp(s)
        sf[self.detect_impossibility(id_fA, nb, True)] = 0
        p = sf / sf.sum()
        omega_f = int(self.rng.choice(range(0, len(p)), 1, p=p)[0])
        f_star, omega_star = int(nb[omega_f // self.N_MH]), omega_f % self.N_MH
        self.udpate_forward_vect(id_fA, f_star, omega_star, max_id)
        proba_forward, star = p[omega_f], lf[omega_f]
        lb = self.compute_all_score_MH(id_fA, V_set, False)
        bad = self.detect_impossibility(id_fA, nb, False)
        target = self.likelihood_t / F_t
        sb = lb / F_t
        mb = sb.max()
        if target <= mb - 10:
            target = mb - 10
        sb[sb <= mb - 10] = mb - 10
        target = target - sb.min()
        sb = sb - sb.min()
        eb = np.exp(sb)
        target = np.exp(target)
        eb[bad] = 0
        proba_backward = target / eb.sum()
        with np.errstate(over="ignore"):
            ratio = np.exp((star + proba_backward - self.likelihood_t - proba_forward) / F_t)
        self._mh_accept(ratio, id_fA, f_star, omega_star, max_id, star)
        return self.likelihood_t, n_contigs, min_len, mean_len, max_len, F_t, self.dist_inter_genome(self.cur)

    def step_mtm(self, id_fA, t=0, n_step=1, dt=0):
        """:2936-3070."""
        (n_contigs, min_len, mean_len, max_len), max_id, V_set, nb = self._mh_prologue(id_fA, dt)
        F_t = self.temperature(t, n_step)
        lf = self.compute_all_score_MH(id_fA, V_set, True)
        s = lf / F_t
        s[s == 0] = -np.inf
        max_fwd = s.max()
        s[s <= max_fwd - 600] = -np.inf
        adapt_fwd = np.exp(s - max_fwd)
        sf = np.copy(adapt_fwd)
        sf[self.detect_impossibility(id_fA, nb, True)] = 0
        p = sf / sf.sum()
        omega_f = int(self.rng.choice(range(0, len(p)), 1, p=p)[0])
        f_star, omega_star = int(nb[omega_f // self.N_MH]), omega_f % self.N_MH
        self.udpate_forward_vect(id_fA, f_star, omega_star, max_id)
        star = lf[omega_f]
        self.return_neighbours(f_star, len(nb))                   # V_set_back (:3004): drawn, never used
        lb = self.compute_all_score_MH(f_star, V_set, False)
        sb = lb / F_t
        sb[sb == 0] = -np.inf
        max_bwd = sb.max()
        sb[sb <= max_bwd - 600] = -np.inf
        adapt_bwd = np.exp(sb - max_bwd)
        with np.errstate(over="ignore"):
            ratio = np.exp(max_fwd - max_bwd) * np.sum(adapt_fwd) / np.sum(adapt_bwd)
        self._mh_accept(ratio, id_fA, f_star, omega_star, max_id, star)
        return self.likelihood_t, n_contigs, min_len, mean_len, max_len, F_t, self.dist_inter_genome(self.cur)

    # ---- :475-541
    def dist_inter_genome(self, g1):
        return dist_inter_genome(g1, self.np_init_prev, self.np_init_next, self.np_init_ori,
                                 self.np_init_orientable, self.id_frags_blacklisted, self.is_repeat,
                                 self.n_new_frags, self.n_frags_4_dist)


def sample_candidate(score, n_tmp, rng, F_t=1.0):
    """Candidate filtering + linear-in-log-likelihood weights + draw (cuda_lib_gl.py:1899-1947).
    Returns (index, margin) where margin is the distance of the uniform draw to the nearest CDF
    boundary (None when no draw was consumed), SURVEY H9."""
    remove = list(range(n_tmp, len(score), n_tmp)) + list(range(n_tmp + 1, len(score), n_tmp))
    id_max = int(np.argmax(score))
    filtered = score - score.min()
    filtered[remove] = 0
    max_score = filtered.max()
    filtered = filtered - (max_score - 30)
    filtered[filtered < 0] = 0
    ok = np.nonzero(filtered > 0)[0]
    sub = filtered[ok]
    with np.errstate(all="ignore"):
        sub = sub / sub.sum()
        sub[sub > 0] = np.power(sub[sub > 0], 1. / F_t)
        sub = sub / sub.sum()
    if len(ok) <= 1:
        return id_max, None
    state = rng.get_state()
    out = int(rng.choice(ok, 1, p=sub)[0])
    probe = np.random.RandomState()
    probe.set_state(state)
    u = probe.random_sample()
    cdf = np.cumsum(sub)
    cdf /= cdf[-1]
    margin = float(np.min(np.abs(cdf[:-1] - u))) if len(cdf) > 1 else None
    return out, margin


def dist_inter_genome(g1, init_prev, init_next, init_ori, init_orientable, blacklisted, is_repeat,
                      n_new_frags, n_frags_4_dist):
    """cuda_lib_gl.py:475-541."""
    d = 3.0 * (n_new_frags - n_frags_4_dist)
    norm = 3.0 * (n_new_frags - n_frags_4_dist)
    black = set(blacklisted)
    for f in range(n_new_frags):
        if f in black or is_repeat[f]:
            continue
        prev_t0, next_t0 = init_prev[f], init_next[f]
        tp, tn = g1["prev"][f], g1["next"][f]
        prev_t1 = g1["id_d"][tp] if tp != -1 else tp
        next_t1 = g1["id_d"][tn] if tn != -1 else tn
        ori_t0, ori_t1 = init_ori[f], g1["ori"][f]
        swap = 1
        if ((prev_t1 == prev_t0) and (next_t1 == next_t0)) or ((prev_t1 == next_t0) and (next_t1 == prev_t0)):
            d -= 1
        if init_orientable[f]:
            if ori_t0 != ori_t1:
                prev_t1, next_t1 = next_t1, prev_t1
                swap = -1
            if prev_t0 == prev_t1:
                if prev_t0 == -1:
                    d -= 1
                elif not init_orientable[prev_t1]:
                    d -= 1
                else:
                    d -= 0.5
                    if init_ori[prev_t0] == swap * g1["ori"][prev_t1]:
                        d -= 0.5
            if next_t0 == next_t1:
                if next_t0 == -1:
                    d -= 1
                elif not init_orientable[next_t1]:
                    d -= 1
                else:
                    d -= 0.5
                    if init_ori[next_t0] == swap * g1["ori"][next_t1]:
                        d -= 0.5
        else:
            if (prev_t1 == prev_t0) or (prev_t1 == next_t0):
                d -= 1
            if (next_t1 == next_t0) or (next_t1 == prev_t0):
                d -= 1
    return d / norm if norm != 0 else 0.0
