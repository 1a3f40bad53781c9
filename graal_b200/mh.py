"""Metropolis-Hastings / multiple-try variants of the sampler: the reference's ``step_metropolis_hastings_s_a``
(cuda_lib_gl.py:2836-2934), ``step_mtm`` (:2936-3070) and what they call -- ``set_jumping_distributions_parameters``
(:2563-2588), the MH candidate set ``all_modifications_metropolis`` (:2651-2657: ``pop_out_pop_in_4_mh`` :735-789,
``split_4_mh`` :791-811, ``paste_4_mh`` :813-839, ``transloc_4_mh`` :957-1013), ``compute_all_score_MH`` (:2615-2649),
``udpate_forward_vect`` (:2808-2834), ``detect_impossibility`` (:3072-3100), ``validate_struct`` (:3102-3129).

Unused by the shipped GUI path (main_gl.py runs step_max_likelihood), restated for completeness of the sampler surface:
host logic in NumPy, every structure operation and every likelihood through the same C-ABI as the main path
(graal_apply_move on slots, graal_full_loglik, graal_delta_loglik).  Mixed into ``graal_b200.sampler.sampler``.

The 13 MH candidates of (fA, fB) scored from a BASE structure (the current genome going forward, the proposed one going
backward): 0 eject fA, 1 flip fA, 2 / 3 insert right of fB (pop_in_3, orientation +1 / -1), 4 / 5 insert left of fB
(pop_in_4), 6 / 7 split at fA (upstream 0 / 1), 8 paste fA-fB when both are contig ends (else a copy), 9..12
translocation (split at fA, split at fB only if fB is the matching contig end, paste; else a copy).

Deviations, both forced: ``list(V_set)`` of a CPython-2 set of NumPy ints has no defined order -- the neighbours are taken
in increasing id order; ``argsort`` of the dense normalised matrix is unstable on ties -- a stable sort is used (the
neighbour SET only differs when the delta-th and (delta+1)-th largest values of a row are equal).
"""
import numpy as np

from ._lib import check

I32, F32 = np.int32, np.float32
N_MH = 13


class MetropolisMixin:
    # ------------------------------------------------------------------ jump sets (cuda_lib_gl.py:2548-2588)
    def set_jumping_distributions_parameters(self, delta):
        """For every bin the ``delta`` bins with the largest contact count normalised by the accu product
        (matrix_normalized = hic_matrix_sub_sampled / (norm_vect_accu^T norm_vect_accu)) and their normalised scores."""
        N = int(self.n_frags)
        r, c, v = (np.asarray(a) for a in self._level_coo)
        keep = r != c
        r, c, v = r[keep].astype(np.int64), c[keep].astype(np.int64), v[keep].astype(F32)
        nv = np.asarray(self.norm_vect_accu, dtype=F32).reshape(-1)
        val = v / (nv[r] * nv[c]).astype(F32)
        rows = np.concatenate([r, c]); cols = np.concatenate([c, r]); vals = np.concatenate([val, val])
        # ascending (value, column) order per row == a stable argsort of the dense row restricted to its non-zeros
        order = np.lexsort((cols, vals, rows))
        rows, cols, vals = rows[order], cols[order], vals[order]
        start = np.searchsorted(rows, np.arange(N), side="left")
        end = np.searchsorted(rows, np.arange(N), side="right")
        self.jump_dictionnary = dict()
        delta = int(delta)
        for i in range(N):
            k = min(end[i] - start[i], delta)
            ids = cols[end[i] - k:end[i]]
            sc = vals[end[i] - k:end[i]]
            if k < delta:                   # zero-valued columns: the largest indices not in the row (stable argsort, self removed)
                have = set(int(a) for a in cols[start[i]:end[i]]); have.add(i)
                pad, j = [], N - 1
                while len(pad) < delta - k and j >= 0:
                    if j not in have:
                        pad.append(j)
                    j -= 1
                ids = np.concatenate([np.array(pad[::-1], dtype=np.int64), ids])
                sc = np.concatenate([np.zeros(len(pad), dtype=F32), sc])
            scores = sc.astype(F32)
            with np.errstate(all="ignore"):
                norm_scores = scores / scores.sum()
            d = self.jump_dictionnary[i] = dict()
            d["proba"] = norm_scores
            d["frags"] = np.array(ids, dtype=I32)
            d["set_frags"] = set(int(x) for x in ids)
            distri = np.zeros(N, dtype=F32)
            distri[ids] = norm_scores
            d["distri"] = distri

    # ------------------------------------------------------------------ structures
    def _mh_slots(self):
        from .sampler import CUR, N_LANES, N_TMP_STRUCT
        base = 1 + N_TMP_STRUCT * N_LANES
        return dict(cur=CUR, pop=base, trans1=base + 1, trans2=base + 2, fwd=base + 3)

    def _mh_base(self, forward):
        s = self._mh_slots()
        return s["cur"] if forward else s["fwd"]

    def all_modifications_metropolis(self, id_fA, id_fB, max_id, forward):
        """cuda_lib_gl.py:2651-2657: the 13 MH candidates of (fA, fB) from the base structure into the collector slots."""
        from .sampler import CAND0
        s = self._mh_slots()
        base = self._mh_base(forward)
        h = self.slot_to_host(base)
        # pop_out_pop_in_4_mh (:735-789), modes 0..5 (the pop-out is repeated per mode in the reference; its result is the same)
        m2 = self.apply_move(base, s["pop"], "POP_OUT", id_fA, max_id=max_id)
        self.apply_move(s["pop"], CAND0 + 0, "COPY", 0)
        self.apply_move(base, CAND0 + 1, "FLIP", id_fA)
        self.apply_move(s["pop"], CAND0 + 2, "POP_IN_3", id_fA, id_fB, 1, m2)
        self.apply_move(s["pop"], CAND0 + 3, "POP_IN_3", id_fA, id_fB, -1, m2)
        self.apply_move(s["pop"], CAND0 + 4, "POP_IN_4", id_fA, id_fB, 1, m2)
        self.apply_move(s["pop"], CAND0 + 5, "POP_IN_4", id_fA, id_fB, -1, m2)
        # split_4_mh (:791-811)
        for up in (0, 1):
            self.apply_move(base, CAND0 + 6 + up, "SPLIT", id_fA, aux=up, max_id=max_id)
        # paste_4_mh (:813-839)
        ext = lambda f: h["prev"][f] == -1 or h["next"][f] == -1
        if ext(id_fA) and ext(id_fB):
            self.apply_move(base, CAND0 + 8, "PASTE", id_fA, id_fB, max_id=max_id)
        else:
            self.apply_move(base, CAND0 + 8, "COPY", 0)
        # transloc_4_mh (:957-1013)
        mode = 0
        for up_a in (0, 1):
            m1 = self.apply_move(base, s["trans1"], "SPLIT", id_fA, aux=up_a, max_id=max_id)
            for up_b in (0, 1):
                ok = (h["next"][id_fB] == -1) if up_b == 0 else (h["prev"][id_fB] == -1)
                if ok:
                    mb = self.apply_move(s["trans1"], s["trans2"], "SPLIT", id_fB, aux=up_b, max_id=m1)
                    self.apply_move(s["trans2"], CAND0 + 9 + mode, "PASTE", id_fA, id_fB, max_id=mb)
                else:
                    self.apply_move(base, CAND0 + 9 + mode, "COPY", 0)
                mode += 1

    def compute_likelihood(self, forward=True):
        """cuda_lib_gl.py:1473-1510: full log-likelihood of the current (forward: the proposed) structure."""
        check(self.lib.graal_full_loglik(self.ctx, self._mh_base(forward), None, self._ptr(self.d_out, 2)))
        return np.float64(self._fetch()[2])

    def compute_all_score_MH(self, id_fA, V_set, forward):
        """cuda_lib_gl.py:2615-2649 + multi_likelihood_4_metropolis (:2659-2806): full likelihood of the base structure plus
        the 13 deltas of every neighbour."""
        from .sampler import CAND0
        list_fB = sorted(int(x) for x in V_set)
        base = self._mh_base(forward)
        score = np.zeros(N_MH * len(list_fB), dtype=np.float64)
        likelihood_t = self.compute_likelihood(forward)
        max_id = int(self.slot_to_host(base)["id_c"].max())
        for x, id_fB in enumerate(list_fB):
            self.all_modifications_metropolis(id_fA, id_fB, max_id, forward)
            check(self.lib.graal_delta_loglik(self.ctx, base, CAND0, N_MH, int(id_fA), int(id_fB), max_id, self._ptr(self.d_out, 16)))
            score[x * N_MH:(x + 1) * N_MH] = self._fetch()[16:16 + N_MH] + likelihood_t
        return score

    def _mh_build_one(self, id_fA, id_fB, mode, max_id, forward):
        """The single candidate `mode` (udpate_forward_vect / validate_struct rebuild it, :2808-2826, 3102-3113)."""
        self.all_modifications_metropolis(id_fA, id_fB, max_id, forward)

    def udpate_forward_vect(self, id_fA, id_fB, id_op, max_id):
        """cuda_lib_gl.py:2808-2834: the sampled candidate becomes the 'forward' structure."""
        from .sampler import CAND0
        self._mh_build_one(id_fA, id_fB, int(id_op), max_id, True)
        self.apply_move(CAND0 + int(id_op), self._mh_slots()["fwd"], "COPY", 0)

    def validate_struct(self, id_fA, id_f_sampled, id_op, max_id):
        """cuda_lib_gl.py:3102-3129: commit the sampled candidate to the current genome."""
        from .sampler import CAND0, CUR
        self._mh_build_one(id_fA, id_f_sampled, int(id_op), max_id, True)
        check(self.lib.graal_commit(self.ctx, CUR, CAND0 + int(id_op)))
        self.init_likelihood()

    def detect_impossibility(self, id_fA, list_neighbours, forward):
        """cuda_lib_gl.py:3072-3100: candidate indices that are no real move (paste / translocation need contig ends)."""
        h = self.slot_to_host(self._mh_base(forward))
        return mh_detect_impossibility(h["prev"], h["next"], id_fA, list_neighbours)

    # ------------------------------------------------------------------ the two steps
    def _mh_prologue(self, id_fA, dt):
        from .sampler import CUR
        h = self.slot_to_host(CUR)
        stats = (len(np.unique(h["id_c"])), h["l_cont"].min(), h["l_cont"].mean(), h["l_cont"].max())
        max_id = self.modify_gl_cuda_buffer(id_fA, dt)
        h = self.slot_to_host(CUR)
        V_set = set(self.jump_dictionnary[id_fA]["set_frags"])
        if h["prev"][id_fA] != -1:
            V_set.add(int(h["prev"][id_fA]))
        if h["next"][id_fA] != -1:
            V_set.add(int(h["next"][id_fA]))
        return stats, int(max_id), V_set, np.array(sorted(V_set), dtype=I32)

    def step_metropolis_hastings_s_a(self, id_fA, t=0, n_step=1, dt=0):
        """cuda_lib_gl.py:2836-2934.  Returns (likelihood_t, n_contigs, min_len, mean_len, max_len, F_t, dist)."""
        (n_contigs, min_len, mean_len, max_len), max_id, V_set, nb = self._mh_prologue(id_fA, dt)
        F_t = self.temperature(t, n_step)
        lf = self.compute_all_score_MH(id_fA, V_set, True)
        omega_f, p_fwd = mh_forward_draw(lf, self.detect_impossibility(id_fA, nb, True), F_t, self.rng, 10)
        f_star, omega_star = int(nb[omega_f // N_MH]), omega_f % N_MH
        self.udpate_forward_vect(id_fA, f_star, omega_star, max_id)
        lb = self.compute_all_score_MH(id_fA, V_set, False)
        ratio = mh_ratio_s_a(lf, p_fwd, omega_f, lb, self.detect_impossibility(id_fA, nb, False), self.likelihood_t, F_t, 10)
        self._mh_accept(ratio, id_fA, f_star, omega_star, max_id, lf[omega_f])
        return self.likelihood_t, n_contigs, min_len, mean_len, max_len, F_t, self.dist_inter_genome(self.gpu_vect_frags)

    def step_mtm(self, id_fA, t=0, n_step=1, dt=0):
        """cuda_lib_gl.py:2936-3070 (multiple-try Metropolis).  Same return tuple."""
        (n_contigs, min_len, mean_len, max_len), max_id, V_set, nb = self._mh_prologue(id_fA, dt)
        F_t = self.temperature(t, n_step)
        lf = self.compute_all_score_MH(id_fA, V_set, True)
        omega_f, adapt_fwd, max_fwd = mtm_forward_draw(lf, self.detect_impossibility(id_fA, nb, True), F_t, self.rng)
        f_star, omega_star = int(nb[omega_f // N_MH]), omega_f % N_MH
        self.udpate_forward_vect(id_fA, f_star, omega_star, max_id)
        self.return_neighbours(f_star, len(nb))                   # V_set_back: drawn (consumes the stream) and unused, :3004
        lb = self.compute_all_score_MH(f_star, V_set, False)
        ratio = mtm_ratio(adapt_fwd, max_fwd, lb, F_t)
        self._mh_accept(ratio, id_fA, f_star, omega_star, max_id, lf[omega_f])
        return self.likelihood_t, n_contigs, min_len, mean_len, max_len, F_t, self.dist_inter_genome(self.gpu_vect_frags)

    def _mh_accept(self, ratio, id_fA, f_star, omega_star, max_id, log_likelihood_star):
        r = np.min([1, ratio])
        if r == 1:
            self.validate_struct(id_fA, f_star, omega_star, max_id)
            self.likelihood_t = log_likelihood_star
        else:
            u = self.rng.rand()
            if r >= u:
                self.validate_struct(id_fA, f_star, omega_star, max_id)
                self.likelihood_t = log_likelihood_star


# ---------------------------------------------------------------------- pure host arithmetic (pinned against the reference lines)
def mh_detect_impossibility(prev, nxt, id_fA, list_neighbours):
    idx_impossibility = []
    is_fA_pastable = prev[id_fA] == -1 or nxt[id_fA] == -1
    for idx, id_fB in enumerate(list_neighbours):
        is_fB_pastable = prev[id_fB] == -1 or nxt[id_fB] == -1
        if not (is_fB_pastable and is_fA_pastable):
            idx_impossibility.append(N_MH * idx + 8)
        if not (nxt[id_fB] == -1):
            idx_impossibility.append(N_MH * idx + 9)
            idx_impossibility.append(N_MH * idx + 11)
        if not (prev[id_fB] == -1):
            idx_impossibility.append(N_MH * idx + 10)
            idx_impossibility.append(N_MH * idx + 12)
    return idx_impossibility


def _choice(rng, n, p):
    return int(rng.choice(range(0, n), 1, p=p)[0])


def mh_forward_draw(log_score_forward, discarded, F_t, rng, thresh_overflow):
    """cuda_lib_gl.py:2867-2883."""
    s = log_score_forward / F_t
    max_score = s.max()
    s[s <= max_score - thresh_overflow] = max_score - thresh_overflow
    s = s - s.min()
    score_forward = np.exp(s)
    score_forward[discarded] = 0
    p = score_forward / score_forward.sum()
    return _choice(rng, len(p), p), p


def mh_ratio_s_a(log_score_forward, p_score_forward, omega_f, log_score_backward, discarded_bwd, likelihood_t, F_t, thresh_overflow):
    """cuda_lib_gl.py:2884-2911."""
    proba_forward = p_score_forward[omega_f]
    log_likelihood_star = log_score_forward[omega_f]
    target_likelihood = likelihood_t / F_t
    s = log_score_backward / F_t
    max_score_back = s.max()
    if target_likelihood <= max_score_back - thresh_overflow:
        target_likelihood = max_score_back - thresh_overflow
    s[s <= max_score_back - thresh_overflow] = max_score_back - thresh_overflow
    target_likelihood = target_likelihood - s.min()
    s = s - s.min()
    score_backward = np.exp(s)
    target_likelihood = np.exp(target_likelihood)
    score_backward[discarded_bwd] = 0
    proba_backward = target_likelihood / score_backward.sum()
    with np.errstate(over="ignore"):
        return np.exp((log_likelihood_star + proba_backward - likelihood_t - proba_forward) / F_t)


def mtm_forward_draw(log_score_forward, discarded, F_t, rng, thresh_overflow=600):
    """cuda_lib_gl.py:2967-3000."""
    s = log_score_forward / F_t
    s[s == 0] = -np.inf
    max_score = s.max()
    s[s <= max_score - thresh_overflow] = -np.inf
    adapt = np.exp(s - max_score)
    score_forward = np.copy(adapt)
    score_forward[discarded] = 0
    p = score_forward / score_forward.sum()
    return _choice(rng, len(p), p), adapt, max_score


def mtm_ratio(adapt_score_fwd, max_forward, log_score_backward, F_t, thresh_overflow=600):
    """cuda_lib_gl.py:3006-3050: exp(max_fwd - max_bwd) * sum(adapted forward) / sum(adapted backward)."""
    s = log_score_backward / F_t
    s[s == 0] = -np.inf
    max_backward = s.max()
    s[s <= max_backward - thresh_overflow] = -np.inf
    adapt_bwd = np.exp(s - max_backward)
    with np.errstate(over="ignore"):
        return np.exp(max_forward - max_backward) * np.sum(adapt_score_fwd) / np.sum(adapt_bwd)
