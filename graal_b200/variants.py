"""The rest of the reference's ``sampler`` surface around the scoring path (cuda_lib_gl.py), as a mixin of
``graal_b200.sampler.sampler``: the per-mode / per-neighbour entry points the step methods are made of, the validation
step that scores every candidate by a FULL likelihood, the older proposal rule and its step, the genome scramblers,
the local inversion move, the parameter packers.  Same names, argument meaning and return values as the reference;
every structure operation and every likelihood goes through the C-ABI (graal_build_candidates, graal_apply_move,
graal_full_loglik, graal_delta_loglik) -- host code here is only the reference's own host logic.

Reference map (cuda_lib_gl.py): update_neighbourhood :717-733, pop_out_pop_in_4_mh :735-789, split_4_mh :791-811,
paste_4_mh :813-839, pop_out_pop_in :841-914, transloc :916-955, transloc_4_mh :957-1013, diagnosis :1016-1042,
new_perform_modificationS :1045-1048, local_flip :1056-1154, setup_rippe_parameters_4_simu :1186-1201,
setup_model_parameters :1216-1227, insert_repeats :1512-1519, modify_genome :1521-1537, return_rippe_vals :1982-1984,
compute_likelihood_4_nuisance :1986-2020, debug_step_max_likelihood :2109-2293, old_return_neighbours :2333-2360,
setup_distri_frags :2363-2390, stream_likelihood :2392-2546, define_neighbourhood :2548-2561,
multi_likelihood_4_metropolis :2659-2806, modify_param_simu :3131-3138, step_max_likelihood_4_visu :3140-3323.

Not restated: simulate_rippe_contacts (:1355-1421, curand data simulation), estimate_parameters_rv (:1296-1352, needs
optim_rippe_curve_update's second model whose kernels -- kernels4.cu -- are not in the reference tree), loadProgram /
load_gl_cuda_* / update_texture_4_sub / display_modif_vect / meminfo (PyCUDA + OpenGL plumbing).
"""
import numpy as np

from ._lib import GraalError, check
from . import rippe as opti

I32, F32 = np.int32, np.float32
N_TMP = 13

PARAM_SIMU_EXP_FIELDS = ("d0", "d1", "d_max", "alpha_0", "alpha_1", "alpha_2", "fact", "v_inter")


def linear_score_draw(score, n_tmp_struct, thresh_overflow, F_t, rng, empty_is_max=False):
    """The candidate draw of debug_step_max_likelihood (:2228-2261, thresh 600, no temperature) and
    step_max_likelihood_4_visu (:3242-3287, thresh 30, temperature): weights are the SHIFTED LOG-likelihoods themselves
    (not their exponentials).  Returns (sample_out, sub_score)."""
    scores_2_remove = list(range(n_tmp_struct, len(score), n_tmp_struct)) + list(range(n_tmp_struct + 1, len(score), n_tmp_struct))
    id_max = score.argmax()
    filtered_score = score - score.min()
    filtered_score[scores_2_remove] = 0
    max_score = filtered_score.max()
    filtered_score = filtered_score - (max_score - thresh_overflow)
    filtered_score[filtered_score < 0] = 0
    id_ok = np.nonzero(filtered_score > 0)[0]
    sub_score = filtered_score[id_ok]
    with np.errstate(all="ignore"):
        sub_score = sub_score / sub_score.sum()
        if F_t is not None:
            sub_score[sub_score > 0] = np.power(sub_score[sub_score > 0], 1. / F_t)
            sub_score = sub_score / sub_score.sum()
    if len(id_ok) == 1 or (empty_is_max and len(id_ok) == 0):
        return int(id_max), sub_score
    return int(rng.choice(id_ok, 1, p=sub_score)[0]), sub_score


def sorted_neighbours_of(level_coo, N, norm_vect_accu=None, drop=()):
    """define_neighbourhood (:2548-2561; norm_vect_accu given) / update_neighbourhood (:717-733; raw counts, the bins of
    ``drop`` removed): per bin the other bins in increasing order of (normalised) contact count -- ``argsort`` of the dense
    row with the bin itself popped.  The dense argsort is unstable on ties; a stable sort is used (zero columns in
    increasing index order first, then the non-zeros by (value, column))."""
    r, c, v = (np.asarray(a) for a in level_coo)
    keep = r != c
    r, c, v = r[keep].astype(np.int64), c[keep].astype(np.int64), v[keep].astype(F32)
    if norm_vect_accu is not None:
        nv = np.asarray(norm_vect_accu, dtype=F32).reshape(-1)
        v = v / (nv[r] * nv[c]).astype(F32)
    rows = np.concatenate([r, c]); cols = np.concatenate([c, r]); vals = np.concatenate([v, v])
    nz = vals != 0
    rows, cols, vals = rows[nz], cols[nz], vals[nz]
    order = np.lexsort((cols, vals, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    start = np.searchsorted(rows, np.arange(N), side="left")
    end = np.searchsorted(rows, np.arange(N), side="right")
    dropped = np.zeros(N, dtype=bool)
    if len(drop):
        dropped[np.asarray(list(drop), dtype=np.int64)] = True
    out = []
    for i in range(N):
        neg = cols[start[i]:end[i]][vals[start[i]:end[i]] < 0]
        pos = cols[start[i]:end[i]][vals[start[i]:end[i]] > 0]
        is_nz = np.zeros(N, dtype=bool)
        is_nz[cols[start[i]:end[i]]] = True
        is_nz[i] = True
        zeros = np.nonzero(~is_nz)[0]
        line = np.concatenate([neg, zeros, pos])
        if len(drop):
            line = line[~dropped[line]]
        out.append(line.astype(I32))
    return out


class VariantsMixin:
    # ------------------------------------------------------------------ candidate builders, one mode / one family at a time
    def pop_out_pop_in(self, id_f_pop, id_f_ins, mode, max_id):
        """cuda_lib_gl.py:841-914: candidate ``mode`` (0..8) of (id_f_pop, id_f_ins) into collector slot ``mode``."""
        mode = int(mode)
        if not 0 <= mode < 9:
            raise GraalError("pop_out_pop_in: mode must be 0..8")
        self.perform_modifications(id_f_pop, id_f_ins, max_id, 1 << mode)

    def transloc(self, id_fA, id_fB, max_id):
        """cuda_lib_gl.py:916-955: the four translocation candidates into collector slots 9..12."""
        self.perform_modifications(id_fA, id_fB, max_id, 0x1E00)

    def new_perform_modificationS(self, id_fA, id_fB, max_id, is_first=True):
        """cuda_lib_gl.py:1045-1048."""
        self.perform_modifications(id_fA, id_fB, max_id, 0x1FFF)

    def pop_out_pop_in_4_mh(self, id_f_pop, id_f_ins, mode, max_id, forward):
        """cuda_lib_gl.py:735-789: MH candidate ``mode`` (0..5) from the current (forward) or the proposed structure."""
        from .sampler import CAND0
        s, base, mode = self._mh_slots(), self._mh_base(forward), int(mode)
        m2 = self.apply_move(base, s["pop"], "POP_OUT", id_f_pop, max_id=max_id)
        if mode == 0:
            self.apply_move(s["pop"], CAND0, "COPY", 0)
        elif mode == 1:
            self.apply_move(base, CAND0 + 1, "FLIP", id_f_pop)
        elif mode in (2, 3):
            self.apply_move(s["pop"], CAND0 + mode, "POP_IN_3", id_f_pop, id_f_ins, 1 if mode == 2 else -1, m2)
        elif mode in (4, 5):
            self.apply_move(s["pop"], CAND0 + mode, "POP_IN_4", id_f_pop, id_f_ins, 1 if mode == 4 else -1, m2)
        else:
            raise GraalError("pop_out_pop_in_4_mh: mode must be 0..5")

    def split_4_mh(self, id_fA, max_id, forward):
        """cuda_lib_gl.py:791-811: MH candidates 6 / 7 (split at fA, upstream 0 / 1)."""
        from .sampler import CAND0
        for up in (0, 1):
            self.apply_move(self._mh_base(forward), CAND0 + 6 + up, "SPLIT", id_fA, aux=up, max_id=max_id)

    def paste_4_mh(self, id_fA, id_fB, max_id, forward):
        """cuda_lib_gl.py:813-839: MH candidate 8 (paste when both bins are contig ends, else a copy)."""
        from .sampler import CAND0
        base = self._mh_base(forward)
        h = self.slot_to_host(base)
        ext = lambda f: h["prev"][f] == -1 or h["next"][f] == -1
        self.apply_move(base, CAND0 + 8, "PASTE" if ext(id_fA) and ext(id_fB) else "COPY", id_fA, id_fB, max_id=max_id)

    def transloc_4_mh(self, id_fA, id_fB, max_id, forward):
        """cuda_lib_gl.py:957-1013: MH candidates 9..12."""
        from .sampler import CAND0
        s, base = self._mh_slots(), self._mh_base(forward)
        h = self.slot_to_host(base)
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

    # ------------------------------------------------------------------ scoring, one neighbour at a time
    def stream_likelihood(self, id_fA, contig_A, len_contig_A, id_fB, id_x, likelihood_t, max_id):
        """cuda_lib_gl.py:2392-2546: the 13 candidates of (fA, fB) and their scores into
        ``self.score[13 * id_x : 13 * id_x + 13]`` (= likelihood_t + delta).  ``contig_A`` / ``len_contig_A`` are
        re-derived on the device (accepted for the signature)."""
        from .sampler import CUR, CAND0
        self.perform_modifications(id_fA, id_fB, max_id)
        check(self.lib.graal_delta_loglik(self.ctx, CUR, CAND0, N_TMP, int(id_fA), int(id_fB), int(max_id), self._ptr(self.d_out, 16)))
        self.score[id_x * N_TMP:(id_x + 1) * N_TMP] = self._fetch()[16:16 + N_TMP] + likelihood_t

    def multi_likelihood_4_metropolis(self, id_fA, contig_A, len_contig_A, id_fB, id_x, gpu_vect_frags, likelihood_t,
                                      likelihood_vect, max_id, score, forward):
        """cuda_lib_gl.py:2659-2806: the 13 MH candidates of (fA, fB) from the current / proposed structure and their
        scores into ``score[13 * id_x : ...]``.  ``gpu_vect_frags`` / ``likelihood_vect`` (the structure and per-pixel
        cache the reference passes) follow from ``forward``."""
        from .sampler import CAND0
        base = self._mh_base(forward)
        self.all_modifications_metropolis(id_fA, id_fB, max_id, forward)
        check(self.lib.graal_delta_loglik(self.ctx, base, CAND0, N_TMP, int(id_fA), int(id_fB), int(max_id), self._ptr(self.d_out, 16)))
        score[id_x * N_TMP:(id_x + 1) * N_TMP] = self._fetch()[16:16 + N_TMP] + likelihood_t

    def compute_likelihood_4_nuisance(self):
        """cuda_lib_gl.py:1986-2020: full likelihood of the current structure under ``self.param_simu_test``."""
        test = getattr(self, "param_simu_test", None)       # filled by step_nuisance_parameters (gpu_param_simu_test, :2088)
        return self.eval_likelihood(test_params=self.param_simu if test is None else test)

    def full_likelihood_of_slot(self, slot):
        check(self.lib.graal_full_loglik(self.ctx, int(slot), None, self._ptr(self.d_out, 2)))
        return np.float64(self._fetch()[2])

    # ------------------------------------------------------------------ proposal rules
    def setup_distri_frags(self):
        """cuda_lib_gl.py:2363-2390: per bin the 10 largest contacts of its level-matrix row and p ~ v^3
        (``distri_frags[i]['xk'] / ['pk']``; the scipy rv_discrete object of the reference is never sampled, :2304)."""
        self.distri_frags = dict()
        for i in range(int(self.n_frags)):
            self.distri_frags[i] = dict(xk=self.distri_xk[i], pk=self.distri_pk[i])
        return self.distri_frags

    def define_neighbourhood(self):
        """cuda_lib_gl.py:2548-2561: ``self.sorted_neighbours[i]`` = the other bins by increasing normalised contact count."""
        self.sorted_neighbours = sorted_neighbours_of(self._level_coo, int(self.n_frags), self.norm_vect_accu)

    def update_neighbourhood(self):
        """cuda_lib_gl.py:717-733: same on the raw counts, rows ``list_frag_to_sample`` only, ``list_to_pop_out`` removed."""
        lines = sorted_neighbours_of(self._level_coo, int(self.n_frags), None, self.list_to_pop_out)
        self.sorted_neighbours = [lines[i] for i in self.list_frag_to_sample]

    def old_return_neighbours(self, id_fA, delta):
        """cuda_lib_gl.py:2333-2360: the ``delta`` (x 15 for a duplicated bin) strongest normalised contacts of the bin,
        each expanded to all its copies, fA's other copies first, blacklisted bins dropped."""
        if getattr(self, "sorted_neighbours", None) is None:
            self.define_neighbourhood()
        ori_id = int(self.h_id_d[id_fA])
        if ori_id in self._dup_set:
            delta = delta * 15
        init_id = np.copy(self.sorted_neighbours[ori_id][-delta:])
        out = []
        if ori_id in self._dup_set:
            d = self.frag_dispatcher[ori_id]
            out.extend(np.setdiff1d(self.collector_id_repeats[d[0]:d[1]], id_fA))
        for id_fB in init_id:
            d = self.frag_dispatcher[id_fB]
            out.extend(self.collector_id_repeats[d[0]:d[1]])
        return [int(e) for e in out if int(e) not in self._black_set]

    # ------------------------------------------------------------------ steps
    def _stats_prologue(self, id_fA, dt):
        """What every step variant does first (:2117-2127, 3148-3163): statistics of the genome BEFORE the relabel."""
        from .sampler import CUR
        h = self.slot_to_host(CUR)
        max_id = self.modify_gl_cuda_buffer(id_fA, dt)
        return int(max_id), len(np.unique(h["id_c"])), h["l_cont"].min(), h["l_cont"].mean(), h["l_cont"].max()

    def debug_step_max_likelihood(self, id_fA, delta, size_block=512, dt=0):
        """cuda_lib_gl.py:2109-2293, the validation step: every candidate of every neighbour is scored by the FULL likelihood
        of its structure (float32 scores), the draw is linear in the shifted scores (threshold 600).  Returns
        (o, n_contigs, min_len, mean_len, max_len, op_sampled, id_f_sampled)."""
        from .sampler import CAND0
        max_id, n_contigs, min_len, mean_len, max_len = self._stats_prologue(id_fA, dt)
        if id_fA in self._black_set:
            return self.o, n_contigs, min_len, mean_len, max_len, -1, id_fA
        self.likelihood_t = self.eval_likelihood()
        id_neighbours = self.return_neighbours(id_fA, delta)
        id_neighbours.sort()
        self.id_neighbours = id_neighbours
        self.score = np.zeros(len(id_neighbours) * N_TMP, dtype=F32)
        for id_x, id_fB in enumerate(id_neighbours):
            self.new_perform_modificationS(id_fA, id_fB, max_id, True)
            for id_mode in range(N_TMP):
                self.score[id_x * N_TMP + id_mode] = self.full_likelihood_of_slot(CAND0 + id_mode)
        or_score = np.copy(self.score)
        sample_out, self.sub_score = linear_score_draw(self.score, N_TMP, 600, None, self.rng)
        id_f_sampled = id_neighbours[sample_out // N_TMP]
        op_sampled = sample_out % N_TMP
        self.test_copy_struct(id_fA, id_f_sampled, op_sampled, max_id)
        self.o = or_score[sample_out]
        return self.o, n_contigs, min_len, mean_len, max_len, op_sampled, id_f_sampled

    def step_max_likelihood_4_visu(self, id_fA, delta, size_block=512, dt=0, t=0, n_step=1):
        """cuda_lib_gl.py:3140-3323: step_max_likelihood with the older proposal rule (old_return_neighbours) and the
        linear draw (threshold 30, temperature).  Returns (o, n_contigs, min_len, mean_len, max_len, op_sampled,
        id_f_sampled, dist, F_t); the reference's own return statement names an undefined ``max_len_bp`` on the scoring
        branch (:3323) -- the max contig length in bins computed beside it is returned."""
        max_id, n_contigs, min_len, mean_len, max_len = self._stats_prologue(id_fA, dt)
        F_t = self.temperature(t, n_step)
        if id_fA not in self._black_set:
            likelihood_t = self.likelihood_t = self.eval_likelihood()
            id_neighbours = self.old_return_neighbours(id_fA, delta)
            id_neighbours.sort()
            self.id_neighbours = id_neighbours
            self.score = np.zeros(len(id_neighbours) * N_TMP, dtype=np.float64)
            for id_x, id_fB in enumerate(id_neighbours):
                self.stream_likelihood(id_fA, None, None, id_fB, id_x, likelihood_t, max_id)
            or_score = np.copy(self.score)
            sample_out, self.sub_score = linear_score_draw(self.score, N_TMP, 30, F_t, self.rng, empty_is_max=True)
            id_f_sampled = id_neighbours[sample_out // N_TMP]
            op_sampled = sample_out % N_TMP
            self.test_copy_struct(id_fA, id_f_sampled, op_sampled, max_id)
            self.o = or_score[sample_out]
        else:
            op_sampled, id_f_sampled = -1, id_fA
        o = self.o
        dist = self.dist_inter_genome(self.gpu_vect_frags)
        self.likelihood_t = o
        return o, n_contigs, min_len, mean_len, max_len, op_sampled, id_f_sampled, dist, F_t

    # ------------------------------------------------------------------ genome scramblers, checks
    def diagnosis(self, c, id_fA, id_fB, id_mut):
        """cuda_lib_gl.py:1016-1042: walk every contig from its first bin along ``next`` and back along ``prev``; the
        reference prints and waits for the operator, here the list of problems is returned (empty = sound)."""
        c.copy_from_gpu()
        problems = []
        for ele in np.nonzero(c.start_bp == 0)[0]:
            len_contig = int(c.l_cont[ele])
            cur_f = int(ele)
            for _ in range(1, len_contig):
                cur_f = int(c.next[cur_f])
            extrem = cur_f
            for _ in range(1, len_contig):
                cur_f = int(c.prev[cur_f])
            if c.circ[ele] == 1 and extrem != c.prev[ele]:
                problems.append(dict(contig=int(c.id_c[ele]), frag=int(ele), id_fA=id_fA, id_fB=id_fB, id_mut=id_mut, kind="circular closure"))
            if cur_f != ele:
                problems.append(dict(contig=int(c.id_c[ele]), frag=int(ele), id_fA=id_fA, id_fB=id_fB, id_mut=id_mut, kind="prev / next walk"))
        return problems

    def _structure_problems(self, c):
        bad = (np.any(c.pos < 0) or np.any(c.l_cont < 0) or np.any(c.l_cont_bp < 0) or np.any(c.start_bp < 0)
               or np.any(c.l_cont_bp - c.start_bp <= 0) or np.any((c.start_bp != 0) * (c.pos == 0))
               or np.any((c.start_bp == 0) * (c.pos != 0)) or np.any(c.next == c.id) or np.any(c.prev == c.id))
        null = np.any(c.l_cont == 0) or np.any(c.l_cont_bp == 0)
        return bool(bad), bool(null)

    def modify_genome(self, n):
        """cuda_lib_gl.py:1521-1537: ``n`` random committed mutations (2n distinct bins, n modes with replacement); the
        reference stops and asks when a structure invariant breaks -- here GraalError."""
        list_breaks = self.rng.choice(int(self.n_new_frags), n * 2, replace=False)
        list_modes = self.rng.choice(self.n_tmp_struct, n, replace=True)
        for i in range(n):
            self.gpu_vect_frags.copy_from_gpu()
            max_id = self.gpu_vect_frags.id_c.max()
            self.test_copy_struct(list_breaks[2 * i], list_breaks[2 * i + 1], list_modes[i], max_id)
            self.gpu_vect_frags.copy_from_gpu()
            bad, null = self._structure_problems(self.gpu_vect_frags)
            if bad or null:
                raise GraalError("modify_genome: structure invariant broken after mutation %d (%d, %d, mode %d)"
                                 % (i, list_breaks[2 * i], list_breaks[2 * i + 1], list_modes[i]))

    def insert_repeats(self, id_f_ins):
        """cuda_lib_gl.py:1512-1519: every duplicated copy is popped out and inserted right of ``id_f_ins`` (mode 7)."""
        for id_ in range(int(self.n_new_frags)):
            self.gpu_vect_frags.copy_from_gpu()
            max_id = self.gpu_vect_frags.id_c.max()
            if self.gpu_vect_frags.rep[id_] == 1:
                self.test_copy_struct(id_, id_f_ins, 7, max_id)

    def local_flip(self, id_fA, mode, max_id):
        """cuda_lib_gl.py:1056-1154: inversion of the window of +-(mode - 11) bins around fA inside its contig, composed
        from the reference's own kernel sequence (pop out every neighbour of the window, re-insert the right-hand ones on
        the left of fA in reverse order and orientation (pop_in_4), the left-hand ones on its right (pop_in_3), flip fA).
        The result is left in collector slot ``mode`` when the reference has one (mode < 13: its collector list is 13 long),
        and returned as a dict of arrays otherwise (the reference's commented-out caller asks for modes 14..16)."""
        from .sampler import CUR, CAND0
        s = self._mh_slots()
        scr, pop, col = s["trans1"], s["pop"], s["trans2"]          # scrambled_gpu_vect_frags, pop_gpu_vect_frags, collector[mode]
        local_delta = int(mode) - 11
        h = self.slot_to_host(CUR)
        pos_fA, id_contig_A, len_contig_A = int(h["pos"][id_fA]), h["id_c"][id_fA], int(h["l_cont"][id_fA])
        neighbours = np.nonzero(h["id_c"] == id_contig_A)[0]
        ordered_neighbours = neighbours[np.argsort(h["pos"][neighbours])]
        orientations_neighbours = h["ori"][ordered_neighbours]
        id_up = max(pos_fA - local_delta, 0)
        id_down = min(pos_fA + local_delta, len_contig_A - 1)
        self.apply_move(CUR, scr, "COPY", 0)
        for i in range(id_up, id_down + 1):
            id_fB = int(ordered_neighbours[i])
            if id_fB != id_fA:
                self.apply_move(scr, pop, "POP_OUT", id_fB, max_id=max_id)
                self.apply_move(pop, scr, "COPY", 0)
                max_id = int(self.slot_to_host(scr)["id_c"].max())
        for j in range(id_down, pos_fA, -1):
            id_fB, ori_fB = int(ordered_neighbours[j]), int(orientations_neighbours[j]) * -1
            self.apply_move(scr, col, "POP_IN_4", id_fB, id_fA, ori_fB, max_id)
            self.apply_move(col, scr, "COPY", 0)
            max_id = int(self.slot_to_host(scr)["id_c"].max())
        for j in range(id_up, pos_fA):
            id_fB, ori_fB = int(ordered_neighbours[j]), int(orientations_neighbours[j]) * -1
            self.apply_move(scr, col, "POP_IN_3", id_fB, id_fA, ori_fB, max_id)
            self.apply_move(col, scr, "COPY", 0)
            max_id = int(self.slot_to_host(scr)["id_c"].max())
        self.apply_move(scr, col, "FLIP", id_fA)
        if 0 <= int(mode) < N_TMP:
            self.apply_move(col, CAND0 + int(mode), "COPY", 0)
        return self.slot_to_host(col)

    # ------------------------------------------------------------------ parameter records
    def setup_rippe_parameters_4_simu(self, kuhn, lm, slope, d, val_inter, d_max):
        """cuda_lib_gl.py:1186-1201: the parameter record with ``fact`` chosen so that the law equals ``val_inter`` at d_max."""
        from .sampler import PARAM_DTYPE, rippe_c1
        kuhn, lm = F32(kuhn), F32(lm)
        c1 = rippe_c1(kuhn, lm, slope)
        slope, d, d_max, val_inter = F32(slope), F32(d), F32(d_max), F32(val_inter)
        param = [kuhn, lm, slope, d]
        # (the reference ran under NumPy 1.x: float32 scalar ** Python float is float64 there, the float32-only factors stay float32)
        rippe = lambda p, dist: (0.53 * (np.float64(p[0]) ** -3.) * np.power((p[1] * dist / p[0]), (p[2])) *
                                 np.exp((p[3] - 2) / ((np.power((p[1] * dist / p[0]), 2) + p[3]))))
        fact = val_inter / rippe(param, d_max)
        return np.array([(kuhn, lm, c1, slope, d, d_max, fact, val_inter)], dtype=PARAM_DTYPE)

    def setup_model_parameters(self, param, d_max):
        """cuda_lib_gl.py:1216-1227: the record of the second (exponential) model; its kernels are not in the reference
        tree, the record is only packed."""
        d0, d1, alpha_0, alpha_1, alpha_2, fact = param
        dt = np.dtype([(k, F32) for k in PARAM_SIMU_EXP_FIELDS], align=True)
        return np.array([(F32(d0), F32(d1), F32(d_max), F32(alpha_0), F32(alpha_1), F32(alpha_2), F32(fact), self.mean_value_trans)], dtype=dt)

    def modify_param_simu(self, param_simu, id_val, val):
        """cuda_lib_gl.py:3131-3138."""
        new_param_simu = np.copy(param_simu)
        if id_val == 0:
            new_param_simu["d"] = F32(val)
        elif id_val == 1:
            new_param_simu["slope"] = F32(val)
        return new_param_simu

    def return_rippe_vals(self, p0):
        """cuda_lib_gl.py:1982-1984."""
        return opti.peval(self.bins, p0)
