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
class OracleSampler:
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
