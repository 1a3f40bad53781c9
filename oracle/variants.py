"""NumPy restatement of the remaining entry points of the reference ``sampler`` (TEST INFRASTRUCTURE -- only tests/,
__graft_entry__.smoke() and bench.py's cpu baseline may import this): the validation step that scores by full
likelihoods, the older proposal rule and its step, local_flip, the scramblers.  Mixed into oracle.sampler.OracleSampler.

Pinned by tests/test_reference_host_logic.py: each method below is compared with the reference's own TEXT
(/root/reference/cuda_lib_gl.py, exec'd under Python 3 with the kernels replaced by oracle.mutations, themselves pinned to
the compiled reference kernels)."""
import numpy as np

from . import mutations as M
from . import likelihood as L

I32, F32 = np.int32, np.float32


def linear_score_draw(score, n_tmp, thresh_overflow, F_t, rng, empty_is_max=False):
    """cuda_lib_gl.py:2228-2261 (debug step: thresh 600, F_t None) / :3242-3287 (4_visu: thresh 30, temperature)."""
    remove = list(range(n_tmp, len(score), n_tmp)) + list(range(n_tmp + 1, len(score), n_tmp))
    id_max = score.argmax()
    f = score - score.min()
    f[remove] = 0
    f = f - (f.max() - thresh_overflow)
    f[f < 0] = 0
    id_ok = np.nonzero(f > 0)[0]
    sub = f[id_ok]
    with np.errstate(all="ignore"):
        sub = sub / sub.sum()
        if F_t is not None:
            sub[sub > 0] = np.power(sub[sub > 0], 1. / F_t)
            sub = sub / sub.sum()
    if len(id_ok) == 1 or (empty_is_max and len(id_ok) == 0):
        return int(id_max), sub
    return int(rng.choice(id_ok, 1, p=sub)[0]), sub


def local_flip(ws, cur, scrambled, collector_mode, id_fA, mode, max_id):
    """cuda_lib_gl.py:1056-1154 on oracle slots: ``scrambled`` = scrambled_gpu_vect_frags, ``collector_mode`` =
    collector_gpu_vect_frags[mode], ``ws.pop`` / ``ws.pop_id_contigs`` = the pop structure.  All persistent."""
    local_delta = mode - 11
    pos_fA, id_contig_A, len_contig_A = cur["pos"][id_fA], cur["id_c"][id_fA], cur["l_cont"][id_fA]
    neighbours = np.nonzero(cur["id_c"] == id_contig_A)[0]
    ordered = neighbours[np.argsort(cur["pos"][neighbours])]
    oris = cur["ori"][ordered]
    id_up = max(pos_fA - local_delta, 0)
    id_down = min(pos_fA + local_delta, len_contig_A - 1)
    M.simple_copy(scrambled, cur)
    for i in range(id_up, id_down + 1):
        id_fB = ordered[i]
        if id_fB != id_fA:
            M.pop_out_frag(ws.pop, scrambled, ws.pop_id_contigs, id_fB, max_id)
            M.simple_copy(scrambled, ws.pop)
            max_id = scrambled["id_c"].max()
    for j in range(id_down, pos_fA, -1):
        M.pop_in_frag_4(collector_mode, scrambled, ordered[j], id_fA, max_id, oris[j] * -1)
        M.simple_copy(scrambled, collector_mode)
        max_id = scrambled["id_c"].max()
    for j in range(id_up, pos_fA):
        M.pop_in_frag_3(collector_mode, scrambled, ordered[j], id_fA, max_id, oris[j] * -1)
        M.simple_copy(scrambled, collector_mode)
        max_id = scrambled["id_c"].max()
    M.flip_frag(collector_mode, scrambled, id_fA)


class OracleVariants:
    # ---- :2548-2561 (dense, the reference way; argsort ties resolved by a stable sort)
    def define_neighbourhood(self):
        nv = np.asarray(self.inp.norm_vect_accu, dtype=F32).reshape(1, -1)
        mat_norm = np.array(nv.T * nv, dtype=F32)
        with np.errstate(all="ignore"):
            self.matrix_normalized = self.hic_matrix_sub_sampled / mat_norm
        tmp_sorted = self.matrix_normalized.argsort(axis=1, kind="stable")
        self.sorted_neighbours = [np.array([x for x in tmp_sorted[i] if x != i], dtype=I32) for i in range(self.n_frags)]

    # ---- :2333-2360
    def old_return_neighbours(self, id_fA, delta):
        if getattr(self, "sorted_neighbours", None) is None:
            self.define_neighbourhood()
        ori_id = self.cur["id_d"][id_fA]
        if ori_id in self.id_frag_duplicated:
            delta = delta * 15
        init_id = np.copy(self.sorted_neighbours[ori_id][-delta:])
        out = []
        if ori_id in self.id_frag_duplicated:
            d = self.dispatcher[ori_id]
            out.extend(np.setdiff1d(self.collector[d[0]:d[1]], id_fA))
        for id_fB in init_id:
            d = self.dispatcher[id_fB]
            out.extend(self.collector[d[0]:d[1]])
        return [int(e) for e in out if e not in self.id_frags_blacklisted]

    def _stats(self):
        c = self.cur
        return len(np.unique(c["id_c"])), c["l_cont"].min(), c["l_cont"].mean(), c["l_cont"].max()

    # ---- :2109-2293
    def debug_step_max_likelihood(self, id_fA, delta, size_block=512, dt=0):
        max_id = self.modify_gl_cuda_buffer(id_fA, dt)
        n_contigs, min_len, mean_len, max_len = self._stats()
        if id_fA in self.id_frags_blacklisted:
            return self.o, n_contigs, min_len, mean_len, max_len, -1, id_fA
        self.likelihood_t = self.eval_likelihood()
        id_neighbours = self.return_neighbours(id_fA, delta)
        id_neighbours.sort()
        self.score = np.zeros(len(id_neighbours) * self.N_TMP, dtype=F32)
        for id_x, id_fB in enumerate(id_neighbours):
            M.perform_modifications(self.ws, self.cur, id_fA, id_fB, max_id)
            for j in range(self.N_TMP):
                self.score[id_x * self.N_TMP + j] = L.evaluate_likelihood(self.ws.collector[j], self.lv, self.param_simu).sum()
        or_score = np.copy(self.score)
        sample_out, self.sub_score = linear_score_draw(self.score, self.N_TMP, 600, None, self.rng)
        id_f_sampled = id_neighbours[sample_out // self.N_TMP]
        op_sampled = sample_out % self.N_TMP
        M.apply_mutation(self.ws, self.cur, id_fA, id_f_sampled, op_sampled, max_id, self.id_contigs)
        self.o = or_score[sample_out]
        return self.o, n_contigs, min_len, mean_len, max_len, op_sampled, id_f_sampled

    # ---- :3140-3323
    def step_max_likelihood_4_visu(self, id_fA, delta, size_block=512, dt=0, t=0, n_step=1):
        max_id = self.modify_gl_cuda_buffer(id_fA, dt)
        n_contigs, min_len, mean_len, max_len = self._stats()
        F_t = self.temperature(t, n_step)
        if id_fA not in self.id_frags_blacklisted:
            likelihood_t = self.likelihood_t = self.eval_likelihood()
            id_neighbours = self.old_return_neighbours(id_fA, delta)
            id_neighbours.sort()
            n = len(id_neighbours)
            self.score = np.zeros(n * self.N_TMP, dtype=np.float64)
            self.delta = np.zeros(n * self.N_TMP, dtype=np.float64)
            for id_x, id_fB in enumerate(id_neighbours):
                self.stream_likelihood(id_fA, id_fB, id_x, likelihood_t, max_id)
            or_score = np.copy(self.score)
            sample_out, self.sub_score = linear_score_draw(self.score, self.N_TMP, 30, F_t, self.rng, empty_is_max=True)
            id_f_sampled = id_neighbours[sample_out // self.N_TMP]
            op_sampled = sample_out % self.N_TMP
            M.apply_mutation(self.ws, self.cur, id_fA, id_f_sampled, op_sampled, max_id, self.id_contigs)
            self.o = or_score[sample_out]
        else:
            op_sampled, id_f_sampled = -1, id_fA
        o = self.o
        dist = self.dist_inter_genome(self.cur)
        self.likelihood_t = o
        return o, n_contigs, min_len, mean_len, max_len, op_sampled, id_f_sampled, dist, F_t

    # ---- :1056-1154
    def local_flip(self, id_fA, mode, max_id):
        if not hasattr(self, "_scrambled"):
            self._scrambled = M.new_slot(self.n_new_frags)
            self._local_collector = M.new_slot(self.n_new_frags)
        local_flip(self.ws, self.cur, self._scrambled, self._local_collector, id_fA, mode, max_id)
        return self._local_collector

    # ---- :1521-1537
    def modify_genome(self, n):
        list_breaks = self.rng.choice(self.n_new_frags, n * 2, replace=False)
        list_modes = self.rng.choice(self.N_TMP, n, replace=True)
        for i in range(n):
            max_id = self.cur["id_c"].max()
            M.apply_mutation(self.ws, self.cur, list_breaks[2 * i], list_breaks[2 * i + 1], list_modes[i], max_id, self.id_contigs)
            assert M.check_invariants(self.cur) == [], (i, M.check_invariants(self.cur))

    # ---- :1512-1519
    def insert_repeats(self, id_f_ins):
        for id_ in range(self.n_new_frags):
            max_id = self.cur["id_c"].max()
            if self.cur["rep"][id_] == 1:
                M.apply_mutation(self.ws, self.cur, id_, id_f_ins, 7, max_id, self.id_contigs)
